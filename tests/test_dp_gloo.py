"""Data-parallel host logic on CPU: 2 processes, gloo backend.  Each rank runs the oracle's
loss/gradient on ITS shard of a global batch, the flat gradient goes through the package's single
all-reduce (dp.allreduce_sum_), and the averaged gradient / AMSGrad step must equal the
single-process full-batch result (the DPCL cost is a mean over the batch, models/dpcl.py:80)."""
import functools
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import models as OM
from oracle import steps as OS

CFG = dict(nb_layers=1, embedding_size=4, window_size=64, hop_size=32)
L, S, B = 1024, 2, 4


def _params():
    return OM.init_separator_params(CFG["window_size"] // 2 + 1, CFG["nb_layers"], 16, CFG["embedding_size"], seed=3)


def _flat_grads(p, batch):
    fn = functools.partial(OS.stft_separator_loss, **CFG)
    tr = OS.trainable(p, ("prediction/",))
    leaves = [v.clone().requires_grad_(True) for v in tr.values()]
    q = dict(p)
    q.update(dict(zip(tr.keys(), leaves)))
    cost, _ = fn(q, *[torch.as_tensor(a) for a in batch])
    grads = torch.autograd.grad(cost, leaves)
    return torch.cat([g.reshape(-1) for g in grads]), float(cost)


def _worker(rank, world_size, port, out):
    import importlib.util
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    # load dp.py on its own: the package __init__ would pull in the CUDA library, which is not needed here
    spec = importlib.util.spec_from_file_location("amss_dp", os.path.join(root, "adaptive-multispeaker-separation_b200", "dp.py"))
    dp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(dp)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    torch.set_num_threads(1)
    batch = OM.synthetic_mixtures(B, S, L, seed=11)
    shard = dp.shard_batch(batch, rank, world_size)
    assert shard[0].shape[0] == B // world_size
    flat, cost = _flat_grads(_params(), shard)
    scale = dp.allreduce_sum_(flat)
    worst = dp.max_over_ranks(float(rank))
    # the same exchange through the per-layer buckets (dp.GradBuckets): two buckets complete "during the backward pass",
    # the third never reports (a layer the loss did not reach) and is flushed by finish()
    flat_b, _ = _flat_grads(_params(), shard)
    n = flat_b.numel()
    cuts = [(0, n // 3, 2), (n // 3, 2 * n // 3, 1), (2 * n // 3, n, 4)]
    gb = dp.GradBuckets(flat_b, cuts)
    gb.ready(1)
    gb.ready(0); gb.ready(0)
    gb.ready(2)                                                    # 1 of 4: stays pending
    assert gb.launched == [True, True, False]
    assert gb.finish() == scale and all(gb.launched)
    assert torch.equal(flat_b, flat)
    clip = 0.5 * float(flat.norm()) * scale                        # a clip that bites: half the mean-gradient norm
    clipped = flat * scale * dp.clip_factor(float(flat.norm()), clip, scale)
    out[rank] = ((flat * scale).numpy(), cost, worst, clipped.numpy(), clip)
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_full_batch_gradient():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    full, cost_full = _flat_grads(_params(), OM.synthetic_mixtures(B, S, L, seed=11))
    g0, c0, w0, k0, clip = out[0]
    g1, c1, w1, k1, _ = out[1]
    assert np.array_equal(g0, g1)                                  # replicas hold identical gradients after the collective
    assert w0 == w1 == 1.0                                         # max-over-ranks helper
    assert abs(0.5 * (c0 + c1) - cost_full) < 1e-5 * abs(cost_full)
    err = np.abs(g0 - full.numpy()).max() / (np.abs(full.numpy()).max() + 1e-30)
    assert err < 1e-4, err
    # gradient clipping under data parallelism: the clipped exchange result equals tf.clip_by_global_norm of the
    # FULL-batch gradient -- norm `clip`, not clip / world_size
    want = full * (clip / max(float(full.norm()), clip))
    assert np.array_equal(k0, k1)
    assert abs(np.linalg.norm(k0) - clip) < 1e-4 * clip
    assert np.abs(k0 - want.numpy()).max() / np.abs(want.numpy()).max() < 1e-4


def test_shard_batch_rejects_ragged_split():
    batch = OM.synthetic_mixtures(3, 2, 256, seed=1)
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("amss_dp", os.path.join(root, "adaptive-multispeaker-separation_b200", "dp.py"))
    dp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(dp)
    try:
        dp.shard_batch(batch, 0, 2)
    except ValueError:
        return
    raise AssertionError("ragged split accepted")
