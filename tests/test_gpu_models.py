"""End-to-end GPU parity: whole training steps and the inference path of the host-side mirror
(Adapt / DPCL / L41Model / KMeans / trainers) against the oracle's restatement of the same
reference graphs, on identical seeded inputs and identical initial parameters."""
import functools

import numpy as np
import pytest
import torch

from oracle import models as M
from oracle import steps as OS
from oracle.kmeans import KMeans as OracleKMeans, random_init_idx

pytestmark = pytest.mark.gpu
REL = 1e-3


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


@pytest.fixture(scope="module")
def amss():
    import amss_b200
    from amss_b200 import models, trainer, ops, synth
    return dict(models=models, trainer=trainer, ops=ops, synth=synth)


def _dev(x):
    return torch.as_tensor(x).cuda().contiguous()


def _copy_params(store, oracle_params):
    for k, v in store.params.items():
        oracle_params[k] = v.detach().cpu().clone()
    return oracle_params


@pytest.mark.parametrize("loss", ["dpcl", "l41"])
def test_stft_separator_train_steps_match_oracle(amss, loss):
    """BASELINE config 1 shape family (STFT + DPCL / L41), reduced: 3 optimisation steps."""
    tr, mo = amss["trainer"], amss["models"]
    B, S, Lw = 2, 2, 4096
    cls = mo.DPCL if loss == "dpcl" else mo.L41Model
    t = tr.STFT_Separator_Trainer(cls, nb_layers=2, layer_size=40, embedding_size=8, learning_rate=1e-3,
                                  window_size=256, hop_size=128, gradient_norm_clip=200.0)
    p = _copy_params(t.store, {})
    fn = functools.partial(OS.stft_separator_loss, nb_layers=2, embedding_size=8, window_size=256, hop_size=128,
                           loss=loss)
    st = OS.Stepper(p, fn, lr=1e-3, clip=200.0)
    for step in range(3):
        mix, nm, I = M.synthetic_mixtures(B, S, Lw, seed=100 + step)
        c_ref, aux = st.step(torch.tensor(mix), torch.tensor(nm), torch.tensor(I))
        c = t.train_step(_dev(mix), _dev(nm), _dev(I))
        assert abs(float(c) - c_ref) < REL * abs(c_ref), (step, float(c), c_ref)
    for k in st.tr:
        assert rel(t.store[k], st.tr[k]) < REL, k


def test_front_separator_train_steps_match_oracle(amss):
    """BASELINE config 2 shape family (frozen adaptive front + DPCL), reduced."""
    tr, mo = amss["trainer"], amss["models"]
    B, S, Lw = 2, 2, 4096
    t = tr.Front_Separator_Trainer(mo.DPCL, nb_layers=2, layer_size=40, embedding_size=8, learning_rate=1e-3,
                                   window_size=64, filters=16, max_pool=32, hop_size=32, with_max_pool=True)
    p = _copy_params(t.store, {})
    fn = functools.partial(OS.front_separator_loss, nb_layers=2, embedding_size=8, max_pool=32, hop=32)
    st = OS.Stepper(p, fn, lr=1e-3)
    front_before = t.store["front/bases/bases"].detach().clone()
    for step in range(3):
        mix, nm, I = M.synthetic_mixtures(B, S, Lw, seed=200 + step)
        c_ref, aux = st.step(torch.tensor(mix), torch.tensor(nm), torch.tensor(I))
        c = t.train_step(_dev(mix), _dev(nm), _dev(I))
        assert abs(float(c) - c_ref) < REL * abs(c_ref), (step, float(c), c_ref)
    for k in st.tr:
        assert rel(t.store[k], st.tr[k]) < REL, k
    assert torch.equal(front_before, t.store["front/bases/bases"].detach())   # front stays frozen


def test_stft_inference_matches_oracle(amss):
    """mixture -> STFT -> embeddings -> k-means masks -> iSTFT (reference call stack 3.4)."""
    tr, mo, ops = amss["trainer"], amss["models"], amss["ops"]
    B, S, Lw = 2, 2, 4096
    inf = tr.STFT_Separator_Inference(mo.DPCL, nb_layers=1, layer_size=24, embedding_size=8, nb_tries=2, nb_steps=4,
                                      window_size=256, hop_size=128, end_assign=True)
    m = inf.model
    p = _copy_params(m.store, {})
    mix, nm, I = M.synthetic_mixtures(B, S, Lw, seed=300)
    pre = M.separator_preprocessing(torch.tensor(mix), torch.tensor(nm), 256, 128, 1.0, 0.0)
    V = M.separator_prediction(p, pre["X"], 1, 8, True)
    Bq, Tt, Fb, E = V.shape
    idx = random_init_idx(B * 2, Tt * Fb, S, np.random.RandomState(3))
    okm = OracleKMeans(S, 2, 4, True, None, 2.0, True)
    _, lab_ref = okm.fit(V.reshape(B, -1, E).detach(), idx)
    # the same stages through the public API
    with torch.no_grad():
        spec, X = ops.stft(_dev(mix), 256, 128)
        Vg = m.prediction(X)
        assert rel(Vg, V) < REL
        sep, lab = m.separate(Vg, X, idx)
        out = m.postprocessing(spec, lab)
    assert torch.equal(m.kmeans.best_try.cpu().long(), okm.last_best)
    # embeddings of a random-init network cluster loosely: a few bins sit on a decision boundary
    assert float((lab.cpu() == lab_ref).float().mean()) > 0.99
    # identical labels in -> identical waveforms out
    sep_ref, _ = M.separate(V, pre["X"], lambda emb: lab.cpu(), S)
    ref = M.postprocessing(sep_ref, pre["stfts"], S, 256, 128)
    assert rel(out, ref) < REL
    assert rel(inf.infer(_dev(mix), init_idx=idx), out) < 1e-6


def test_front_inference_round_trip(amss):
    """front -> separate -> back runs through the public API and returns [B,S,L] waveforms."""
    tr, mo = amss["trainer"], amss["models"]
    inf = tr.Front_Separator_Inference(mo.DPCL, nb_layers=1, layer_size=24, embedding_size=8, nb_tries=2, nb_steps=3,
                                       window_size=64, filters=16, max_pool=32, hop_size=32, with_max_pool=True,
                                       end_assign=True)
    mix, nm, I = M.synthetic_mixtures(2, 2, 2048, seed=301)
    out = inf.infer(_dev(mix), _dev(nm))
    assert out.shape == (2, 2, 2048) and bool(torch.isfinite(out).all())
    # hard masks partition the mixture's front coefficients and the back end is linear in them:
    # the separated sources add up to the reconstruction of the un-masked coefficients
    B = 2
    with torch.no_grad():
        y, am = inf.model.front(_dev(mix), _dev(nm))
        vals = torch.stack([y[:B], torch.zeros_like(y[:B])], 1).reshape(2 * B, *y.shape[1:]).contiguous()
        full = inf.model.back(vals, am, B, 2048)
    assert rel(out.sum(1), full.sum(1)) < 1e-3
