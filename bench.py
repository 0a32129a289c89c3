#!/usr/bin/env python
"""bench.py -- mixtures/sec of the hot path on synthetic LibriSpeech-shaped mixtures (4 s @ 16 kHz, L = 64000).

  python bench.py [--config 1..5] [--gpus N] [--steps K] [--warmup W] [--batch B_per_gpu] [--precision fp32|bf16]
                  [--seconds S]       # time ceil(S / ms_per_step) steps instead of K: a SUSTAINED number
  python bench.py --impl reference ...    # the CPU restatement of the reference graph on the host cores

--config (BASELINE.json `configs`; default 2 = the configuration the headline metric is quoted on):
  1  STFT (512/256) + DPCL, 2 x BLSTM-300, E = 40, 2 speakers, batch 4: training step (the reference's CPU-runnable case)
  2  adaptive front (W = 1024, 256 filters, max_pool 256, frozen) + DPCL 3 x BLSTM-600, E = 40, bf16: training step
  3  STFT + L41 4 x BLSTM-600 (frozen) -> k-means masks -> enhance layer 3 x BLSTM-600 (trained, PIT-L2), batch 32 / GPU
  4  Adapt pre-training (sdr+l2, --beta 0.01, separation mask): analysis fwd/bwd + sparse synthesis fwd/bwd + losses
  5  inference, 3 speakers: STFT -> DPCL 3 x BLSTM-600 -> k-means K = 3 (10 tries x 10 steps) -> masked inverse STFT,
     streaming batches of 64 mixtures in, separated waveforms out

One JSON line on stdout (rank 0).  `value` = device-resident throughput, `e2e` = through the public trainer / inference
API from pinned host buffers incl. the H2D of every batch and the D2H of its result, `roofline` = the dominant kernel of
the config timed with CUDA events around its C-ABI call in host-launched steps of the same process, `kernels` = the same
for the other entry points of the step, `blstm_tc_util_pct` = BLSTM-path flops / time of its kernels / tensor peak,
`cpu_baseline` (+ `parity`) = the oracle port timed (and used as the checker) on this box's host cores.
"""
import argparse
import functools
import json
import math
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

L_SAMPLES = 64000
METRIC = "mixtures/sec (4s, 16kHz, 2-spk) fwd+bwd"
METRIC_INFER = "mixtures/sec (4s, 16kHz, 3-spk) separated (inference)"

CONFIGS = {
    1: dict(kind="stft_train", S=2, batch=4, graph=True,
            model=dict(nb_speakers=2, nb_layers=2, layer_size=300, embedding_size=40, window_size=512, hop_size=256,
                       learning_rate=1e-3),
            workload="STFT (512/256) + DPCL 2xBLSTM-300 E=40, 2-spk, L=64000 (4 s @ 16 kHz), fwd+bwd+AMSGrad"),
    2: dict(kind="front_train", S=2, batch=128, graph=True,
            model=dict(nb_speakers=2, nb_layers=3, layer_size=600, embedding_size=40, window_size=1024, filters=256,
                       max_pool=256, hop_size=256, with_max_pool=True, learning_rate=1e-3),
            workload="adapt front (W=1024, 256 filters, max_pool 256, hop 256, frozen) + DPCL 3xBLSTM-600 E=40, 2-spk, "
                     "L=64000 (4 s @ 16 kHz), fwd+bwd+AMSGrad"),
    3: dict(kind="enhance_train", S=2, batch=32, graph=False,
            model=dict(nb_speakers=2, nb_layers=4, layer_size=600, embedding_size=40, window_size=512, hop_size=256,
                       nb_layers_enhance=3, layer_size_enhance=600, nonlinearity="softmax", nb_tries=10, nb_steps=10,
                       learning_rate=1e-3),
            workload="STFT + L41 4xBLSTM-600 E=40 (frozen) -> k-means masks (10 tries x 10 steps) -> enhance layer "
                     "3xBLSTM-600 (trained, PIT-L2 cost), 2-spk, L=64000, fwd+bwd+AMSGrad"),
    4: dict(kind="adapt_pretrain", S=2, batch=32, graph=True,
            model=dict(nb_speakers=2, window_size=1024, filters=256, max_pool=256, hop_size=256, with_max_pool=True,
                       loss="sdr+l2", separation="mask", beta=0.01, learning_rate=1e-3),
            workload="Adapt pre-training (W=1024, 256 filters, max_pool 256; loss sdr+l2, beta 0.01, separation mask): "
                     "analysis + sparse synthesis autoencoder, 2-spk, L=64000, fwd+bwd+AMSGrad"),
    5: dict(kind="stft_infer", S=3, batch=64, graph=False,
            model=dict(nb_speakers=3, nb_layers=3, layer_size=600, embedding_size=40, window_size=512, hop_size=256,
                       nb_tries=10, nb_steps=10),
            workload="inference, 3-spk: STFT -> DPCL 3xBLSTM-600 E=40 -> k-means K=3 (10 tries x 10 steps, hard) -> masked "
                     "inverse STFT, streaming batches, L=64000; waveforms out"),
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]),
                        tf_sust=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
        except Exception:
            pass
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, source="fallback")


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc = gpu_index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, pw, reasons, capped = [], [], [], set(), 0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            try:
                pw.append(float(f[3]))
            except ValueError:
                pass
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
                    capped += n == "sw_power_cap"
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "sm_min_mhz": min(sm) if sm else None, "power_w_max": max(pw) if pw else None,
                "power_w_median": statistics.median(pw) if pw else None,
                "sw_power_cap_share": capped / len(sm) if sm else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference graph, timed on the host cores (and used as the checker)
# ------------------------------------------------------------------------------------------------
def oracle_setup(cfg_id, params=None):
    """(params, loss_fn(p, mix, nm, I) -> (cost, aux), train_prefixes) of the config's step in the oracle."""
    from oracle import models as OM
    from oracle import steps as OS
    c = CONFIGS[cfg_id]
    m = c["model"]
    if c["kind"] == "front_train":
        p = params or {**OM.init_adapt_params(m["window_size"], m["filters"]),
                       **OM.init_separator_params(m["filters"], m["nb_layers"], m["layer_size"], m["embedding_size"])}
        fn = functools.partial(OS.front_separator_loss, nb_layers=m["nb_layers"], embedding_size=m["embedding_size"],
                               max_pool=m["max_pool"], hop=m["hop_size"])
        return p, fn, ("prediction/",)
    if c["kind"] == "stft_train":
        p = params or OM.init_separator_params(m["window_size"] // 2 + 1, m["nb_layers"], m["layer_size"], m["embedding_size"])
        fn = functools.partial(OS.stft_separator_loss, nb_layers=m["nb_layers"], embedding_size=m["embedding_size"],
                               window_size=m["window_size"], hop_size=m["hop_size"])
        return p, fn, ("prediction/",)
    if c["kind"] == "adapt_pretrain":
        p = params or OM.init_adapt_params(m["window_size"], m["filters"])

        def fn(p, mix, nm, I):
            return OM.adapt_pretraining_cost(p, mix, nm, max_pool=m["max_pool"], hop=m["hop_size"], loss=m["loss"],
                                             separation=m["separation"], beta=m["beta"])
        return p, fn, ("front/", "back/")
    if c["kind"] == "enhance_train":
        p = params or {**OM.init_separator_params(m["window_size"] // 2 + 1, m["nb_layers"], m["layer_size"],
                                                  m["embedding_size"], with_speaker_vectors=True),
                       **OM.init_enhance_params(m["window_size"] // 2 + 1, m["nb_layers_enhance"], m["layer_size_enhance"])}
        fn = functools.partial(OS.stft_enhance_loss, nb_layers=m["nb_layers"], embedding_size=m["embedding_size"],
                               window_size=m["window_size"], hop_size=m["hop_size"], nb_layers_enhance=m["nb_layers_enhance"],
                               nb_tries=m["nb_tries"], nb_steps=m["nb_steps"])
        return p, fn, ("enhance/",)
    if c["kind"] == "stft_infer":
        p = params or OM.init_separator_params(m["window_size"] // 2 + 1, m["nb_layers"], m["layer_size"], m["embedding_size"])
        fn = functools.partial(OS.stft_inference, nb_layers=m["nb_layers"], embedding_size=m["embedding_size"],
                               window_size=m["window_size"], hop_size=m["hop_size"], S=c["S"], nb_tries=m["nb_tries"],
                               nb_steps=m["nb_steps"])
        return p, fn, ()
    raise ValueError(cfg_id)


def cpu_steps(cfg_id, batch, steps, warmup, threads=None):
    """Times the oracle's step of the config on `threads` host threads: (mixtures/s, s/step, threads)."""
    import torch
    from oracle import models as OM
    from oracle import steps as OS
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    c = CONFIGS[cfg_id]
    p, fn, prefixes = oracle_setup(cfg_id)
    mix, nm, I = OM.synthetic_mixtures(batch, c["S"], L_SAMPLES, seed=42)
    mix, nm, I = torch.tensor(mix), torch.tensor(nm), torch.tensor(I)
    if c["kind"] == "stft_infer":
        with torch.no_grad():
            run = lambda: fn(p, mix, nm, I)  # noqa: E731
            for _ in range(warmup):
                run()
            t0 = time.time()
            for _ in range(steps):
                run()
    else:
        st = OS.Stepper(p, fn, train_prefixes=prefixes, lr=c["model"]["learning_rate"])
        for _ in range(warmup):
            st.step(mix, nm, I)
        t0 = time.time()
        for _ in range(steps):
            st.step(mix, nm, I)
    dt = time.time() - t0
    return batch * steps / dt, dt / steps, threads


CPU_SAMPLE = {1: (4, 6, 1), 2: (1, 8, 1), 3: (1, 2, 1), 4: (1, 6, 1), 5: (1, 2, 0)}     # (batch, steps, warm-up) of the in-run leg
CPU_NOTE = "torch-CPU restatement of the TF graph (TF 1.x / Python 2 not installable)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cid = args.config
    heavy = cid in (3, 5)
    steps, warmup = max(1, min(args.steps, 3 if heavy else 10)), max(0, min(args.warmup, 1 if heavy else 2))
    batch = {1: 4, 2: 2, 3: 1, 4: 2, 5: 1}[cid]
    v, per_step, threads = cpu_steps(cid, batch, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC if CONFIGS[cid]["kind"] != "stft_infer" else METRIC_INFER, "value": v, "unit": "mixtures/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": CONFIGS[cid]["workload"], "batch_per_step": batch, "baseline_config": cid},
        "cpu_baseline": {"value": v, "unit": "mixtures/s", "cores": threads, "kind": "port",
                         "sample": f"{steps} step(s) of batch {batch} after {warmup} warm-up; {CPU_NOTE}"},
        "e2e": {"value": v, "unit": "mixtures/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# roofline formulas per C-ABI entry point: args are the positional arguments of the call
# ------------------------------------------------------------------------------------------------
def _rl_analysis_mix(a):      # (x, filt, B, S, L, W, N, pool, hop, mode, prec, ...)
    B, S, L, W, N = a[2:7]
    return "tensor", 2.0 * L * W * N * B * (S + 1), "filterbank analysis: conv SAME stride 1 + max-pool/arg-max fused (adapt.py:115-117)"


def _rl_analysis(a):          # (x, filt, Bt, L, W, N, pool, hop, mode, prec, ...)
    Bt, L, W, N, pool, hop, mode = a[2:9]
    T = L if mode != 2 else -(-L // hop)
    return "tensor", 2.0 * T * W * N * Bt, "filterbank analysis (adapt.py:115-122)"


def _rl_analysis_bwd(a):      # (x, dy, am, Bt, L, W, N, Tp, ...)
    Bt, L, W, N, Tp = a[3:8]
    return ("hbm", Bt * (Tp * N * 12.0 + L * 4.0) + W * N * 4.0,
            "filter gradient through the max-pool arg-max: dy + int64 arg-max + waveform in, dfilt out", {"fma": 1.0 * Bt * Tp * N * W, "fma_per_lds": 1})


def _rl_synth_fwd(a):         # (vals, am, filt2, B, S, L, W, N, Tp, pool, hop, out, ...)
    B, S, L, W, N, Tp = a[3:9]
    return ("hbm", B * S * Tp * N * 4.0 + B * Tp * N * 8.0 + W * N * 4.0 + B * S * L * 4.0,
            "sparse overlap-add synthesis: unpool + conv2d_transpose fused (adapt.py:205-252)", {"fma": 1.0 * B * S * Tp * N * W, "fma_per_lds": S})


def _rl_synth_bwd(a):         # (dout, vals, am, filt2, B, S, L, W, N, Tp, ...)
    B, S, L, W, N, Tp = a[4:10]
    return ("hbm", B * S * L * 4.0 + 2.0 * B * S * Tp * N * 4.0 + B * Tp * N * 8.0 + 2.0 * W * N * 4.0, "synthesis backward: dvals + dfilt2",
            {"fma": 2.0 * B * S * Tp * N * W, "fma_per_lds": 1})


def _rl_blstm_fwd(a):         # (x, kf, bf, kb, bb, B, T, I, H, ...)
    B, T, I, H = a[5:9]
    return "tensor", 2.0 * T * B * 2 * (I + H) * 4 * H, "BLSTM forward: input projection + recurrence (utils/ops.py:358-383)"


def _rl_blstm_bwd(a):         # (x, kf, kb, y, dy, saved, B, T, I, H, ...)
    B, T, I, H = a[6:10]
    return "tensor", 2.0 * 2.0 * T * B * 2 * (I + H) * 4 * H, "BLSTM backward: BPTT recurrence + dW + dx"


def _rl_gemm_bf16(a):         # (A, lda, a_mn, B, ldb, b_mn, bias, M, N, K, ...)
    M, N, K = a[7:10]
    return "tensor", 2.0 * M * N * K, "GEMM (Conv1D k=1 head / its gradients, utils/ops.py:486-503)"


def _rl_gemm(a):              # (A, lda, B, ldb, bias, M, N, K, ...)
    M, N, K = a[5:8]
    return "tensor", 2.0 * M * N * K, "GEMM"


def _rl_kmeans(a):            # (X, init, ns, B, L, E, K, tries, iters, ...)
    B, L, E, K, tries, iters = a[3:9]
    return "hbm", float(B) * tries * (iters + 2) * L * E * 4.0, "k-means fit: (iters+2) passes over X per try (Kmeans_2.py:86-188)"


def _rl_stft(a):              # (x, R, L, frame, hop, spec, mag)
    R, L, frame, hop = a[1:5]
    T, F = 1 + (L - frame) // hop, frame // 2 + 1
    return "hbm", R * (L * 4.0 + T * F * ((8.0 if a[5] else 0.0) + (4.0 if a[6] else 0.0))), "STFT: frame + hann + rFFT + |.| fused (network.py:480-492)"


def _rl_stft_labels(a):       # (non_mix, B, S, L, frame, hop, labels, mag)
    B, S, L, frame, hop = a[1:6]
    T, F = 1 + (L - frame) // hop, frame // 2 + 1
    return "hbm", B * (S * L * 4.0 + T * F * (1.0 + (4.0 * S if a[7] else 0.0))), "STFT of the sources + arg-max labels (network.py:489-502)"


def _rl_istft(a):             # (spec, labels, masks, B, S, T, frame, hop, out)
    B, S, T, frame, hop = a[3:8]
    F = frame // 2 + 1
    return "hbm", B * (T * F * (8.0 + (4.0 if a[1] else 4.0 * S)) + S * ((T - 1) * hop + frame) * 4.0), "masked inverse STFT + overlap-add (network.py:584-607)"


def _rl_dpcl(a, passes):      # (V, labels, B, TF, E, ...)
    return None


ROOFLINES = {
    "amss_filterbank_analysis_mix_fwd": _rl_analysis_mix, "amss_filterbank_analysis_fwd": _rl_analysis,
    "amss_filterbank_analysis_bwd": _rl_analysis_bwd, "amss_filterbank_synthesis_fwd": _rl_synth_fwd,
    "amss_filterbank_synthesis_bwd": _rl_synth_bwd, "amss_blstm_fwd": _rl_blstm_fwd, "amss_blstm_bwd": _rl_blstm_bwd,
    "amss_gemm_bf16": _rl_gemm_bf16, "amss_gemm": _rl_gemm, "amss_kmeans_fit": _rl_kmeans, "amss_stft_fwd": _rl_stft,
    "amss_stft_labels": _rl_stft_labels, "amss_istft_masked_fwd": _rl_istft,
}
BLSTM_PATH = ("amss_blstm_fwd", "amss_blstm_bwd", "amss_gemm_bf16", "amss_gemm")


def kernel_table(timed, nsteps, step_ms, pk, sustained):
    """Per entry point: ms per step, share, algorithmic work per step, achieved rate, fraction of the measured peak."""
    rows = []
    tf_peak = pk["tf_sust"] if sustained else pk["tf_burst"]
    for name, calls in timed.items():
        ms = sum(m for _, m in calls) / nsteps
        row = {"entry": name, "calls_per_step": len(calls) / nsteps, "ms_per_step": ms, "share_of_step": ms / step_ms}
        fn = ROOFLINES.get(name)
        if fn is not None:
            work, bound, what, fma, per_lds = 0.0, None, None, 0.0, 1
            for a, _ in calls:
                spec = fn(a)
                bound, w, what = spec[:3]
                work += w
                if len(spec) > 3:
                    fma += spec[3]["fma"]
                    per_lds = spec[3]["fma_per_lds"]
            work /= nsteps
            if bound == "tensor":
                ach = work / (ms / 1e3) / 1e12
                row.update(bound="tensor", achieved=ach, peak=tf_peak, unit="TFLOP/s", frac=ach / tf_peak, work_per_step=work)
            else:
                ach = work / (ms / 1e3) / 1e9
                row.update(bound="hbm", achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"], work_per_step=work)
            if fma > 0:
                # sparse CUDA-core kernels: every multiply-add of an atom reads one filter tap / signal sample from shared
                # memory (shared across the S rows of a mixture in the forward): the limiter is the 128 B/clk/SM shared-memory
                # pipe, not HBM -- the algorithmic bytes above are a few MB per step
                roof = 148 * 32 * 1.965e9 * per_lds / 1e12
                tf = fma / nsteps / (ms / 1e3) / 1e12
                row.update(limiter="shared-memory loads (1 per %d multiply-adds)" % per_lds, achieved_tfma=tf, smem_roof_tfma=roof,
                           frac_of_smem_roof=tf / roof)
            row["what"] = what
        rows.append(row)
    rows.sort(key=lambda r: -r["ms_per_step"])
    return rows


# ------------------------------------------------------------------------------------------------
# parity of the benchmarked path against the oracle, at the bench geometry (runs inside the cpu_baseline leg)
# ------------------------------------------------------------------------------------------------
def parity_vs_oracle(cfg_id, precision, n_mix=2, seed=4242, kmeans=True):
    """Runs ONE forward of config 1 / 2 at full geometry (L = 64000) on `n_mix` mixtures through the package (given
    precision) and through the oracle with identical parameters and inputs; returns the error budget of the outputs
    north_star names: loss, embeddings V, front arg-max / labels, k-means masks.  The oracle is the checker only."""
    import numpy as np
    import torch
    from amss_b200 import models, trainer, ops
    from oracle import models as OM
    c = CONFIGS[cfg_id]
    m = c["model"]
    S = c["S"]
    mix, nm, I = OM.synthetic_mixtures(n_mix, S, L_SAMPLES, seed=seed)
    dev = [torch.as_tensor(a).cuda() for a in (mix, nm, I)]
    cpu = [torch.tensor(a) for a in (mix, nm, I)]
    out = {"config": cfg_id, "precision": precision, "mixtures": n_mix}
    if c["kind"] == "front_train":
        t = trainer.Front_Separator_Trainer(models.DPCL, precision=precision, **m)
    else:
        t = trainer.STFT_Separator_Trainer(models.DPCL, precision=precision, **m)
    p = {k: v.detach().cpu().clone() for k, v in t.store.params.items()}
    _, fn, _ = oracle_setup(cfg_id, p)
    with torch.no_grad():
        cost_ref, aux = fn(p, *cpu)
        if c["kind"] == "front_train":
            y, am = t.model.front(dev[0], dev[1])
            inp = t.sepNet.plugged_inputs(y, n_mix)
            V = t.sepNet.prediction(inp["X"].contiguous())
            cost = t.sepNet.cost(V, inp["labels"], dev[2])
            fr = aux["front"]
            out["front_y_rel_max"] = float((y.cpu() - fr["y"]).abs().max() / fr["y"].abs().max())
            out["front_argmax_agree"] = float((am.cpu() == fr["argmax"]).float().mean())
            lab_ref = aux["inp"]["argmax"].reshape(n_mix, -1)
            out["labels_agree"] = float((inp["labels"].reshape(n_mix, -1).cpu().long() == lab_ref).float().mean())
            Xin = inp["X"]
        else:
            pre = t.model.preprocessing(dev[0], dev[1])
            V = t.model.prediction(pre["X"])
            cost = t.model.cost(V, pre["labels"], dev[2])
            out["stft_mag_rel_max"] = float((pre["X"].cpu() - aux["pre"]["X"]).abs().max() / aux["pre"]["X"].abs().max())
            lab_ref = aux["pre"]["argmax"].reshape(n_mix, -1)
            out["labels_agree"] = float((pre["labels"].reshape(n_mix, -1).cpu().long() == lab_ref).float().mean())
            Xin = pre["X"]
        Vr = aux["V"]
        d = (V.cpu() - Vr).double()
        out["loss"], out["loss_oracle"] = float(cost), float(cost_ref)
        out["loss_rel"] = abs(float(cost) - float(cost_ref)) / abs(float(cost_ref))
        out["V_rel_max"] = float(d.abs().max() / Vr.abs().max())
        out["V_rel_rms"] = float(d.pow(2).mean().sqrt() / Vr.double().pow(2).mean().sqrt())
        out["V_cosine_min"] = float((V.cpu().double() * Vr.double()).sum(-1).min())
        if kmeans:
            # masks: the package's k-means on its own embeddings vs the same kernel on the oracle's embeddings, same initial
            # rows (hard labels; agreement up to nothing -- identical initial rows give identical cluster ids)
            Bq, Tt, Fb, E = V.shape
            rng = np.random.RandomState(7)
            init = np.stack([rng.choice(Tt * Fb, size=S, replace=False) for _ in range(n_mix * 10)]).astype(np.int32)
            idx = torch.as_tensor(init).cuda()
            _, lab_a, _, _ = ops.kmeans_fit(V.reshape(Bq, Tt * Fb, E).contiguous(), idx, S, 10, 10)
            _, lab_b, _, _ = ops.kmeans_fit(Vr.reshape(Bq, Tt * Fb, E).cuda().contiguous(), idx, S, 10, 10)
            agree = (lab_a == lab_b).float().mean(1)
            out["kmeans_mask_agree"] = float(torch.maximum(agree, 1 - agree).mean() if S == 2 else agree.mean())
            w = Xin.reshape(Bq, -1).abs()
            hit = ((lab_a == lab_b).float() * w).sum(1) / w.sum(1)
            out["kmeans_mask_agree_energy_weighted"] = float(torch.maximum(hit, 1 - hit).mean() if S == 2 else hit.mean())
    del t
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    if args.stock_front:
        os.environ["AMSS_NO_LINEAR_MIX"] = "1"
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import amss_b200  # noqa: F401
    from amss_b200 import models, trainer, synth, _lib, dp
    numa_node = None
    if world > 1:           # one process per GPU: keep its host buffers and its launcher thread next to the GPU
        try:
            pr = torch.cuda.get_device_properties(local)
            numa_node = dp.bind_to_gpu_numa_node(f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0")
        except Exception:   # noqa: BLE001 -- host-side plumbing, never fatal
            numa_node = None

    cid = args.config
    c = CONFIGS[cid]
    m = dict(c["model"])
    kind, S = c["kind"], c["S"]
    B = args.batch or c["batch"]
    precision = args.precision

    # ---- the workload ------------------------------------------------------------------------------------------------
    if kind == "front_train":
        t = trainer.Front_Separator_Trainer(models.DPCL, precision=precision, **m)
    elif kind == "stft_train":
        t = trainer.STFT_Separator_Trainer(models.DPCL, precision=precision, **m)
    elif kind == "enhance_train":
        t = trainer.STFT_Separator_enhance_Trainer(models.L41Model, precision=precision, **m)
    elif kind == "adapt_pretrain":
        t = trainer.Adapt_Pretrainer(precision=precision, **m)
    else:
        t = trainer.STFT_Separator_Inference(models.DPCL, precision=precision, **m)
    training = kind != "stft_infer"
    if training and args.flat_allreduce:
        t.args["overlap_allreduce"] = False
    use_graph = bool(args.cuda_graph and c["graph"] and precision == "bf16" and args.warmup >= 3 and training)
    if use_graph:
        t.enable_cuda_graph()
    stream = synth.SyntheticStream(B, S, L_SAMPLES, seed=42, rank=rank, pool=2)
    host_batches = [next(stream) for _ in range(2)]
    if not training:
        feed_batches = [(hb[0], None, None) for hb in host_batches]       # inference reads the mixtures only
    elif args.ship_mix:
        feed_batches = host_batches
    else:
        feed_batches = [(None, hb[1], hb[2]) for hb in host_batches]      # the mixture is built on the device from the sources
    dev_batches = [[None if a is None else torch.as_tensor(a).cuda() for a in hb] for hb in feed_batches]

    if training:
        def step(i):
            return t.train_step(*dev_batches[i % 2])
    else:
        def step(i):
            return t.infer(dev_batches[i % 2][0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # ---- warm-up -----------------------------------------------------------------------------------------------------
    for i in range(args.warmup):
        step(i)
    barrier()
    steps = args.steps
    if args.seconds:
        probe = timed(step, 3) / 3
        steps = max(args.steps, int(math.ceil(args.seconds * 1e3 / probe)))

    # ---- device-resident timing ----------------------------------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count()
    g0 = t.graph_kernel_launches() if training else 0
    ms = timed(step, steps)
    launches = (_lib.launch_count() - n0) + ((t.graph_kernel_launches() - g0) if training else 0)

    # ---- end to end through the public API: every step copies ITS batch from pinned host memory (side stream, one step
    #      ahead) and reads ITS result back (the loss, one step delayed; for inference the separated waveforms) -----------
    pinned_batches = [trainer.DevicePrefetcher.pin(hb) for hb in feed_batches]

    class Feed:
        def __init__(self):
            self.i = 0

        def __iter__(self):
            return self

        def __next__(self):
            b = pinned_batches[self.i % 2]
            self.i += 1
            return b

    results = []
    e2e_steps = steps if not args.seconds else max(args.steps, steps // 4)
    if training:
        t.train(Feed(), min(2, args.warmup) + 1)
        run_e2e = lambda: results.extend(t.train(Feed(), e2e_steps))  # noqa: E731
    else:
        for _ in t.inference(Feed(), 2):
            pass
        run_e2e = lambda: results.extend(float(o[0, 0, 1000]) for o in t.inference(Feed(), e2e_steps))  # noqa: E731
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_e2e()
    e1.record()
    barrier()
    ms_e2e_t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms_e2e_t, op=dist.ReduceOp.MAX)
    ms_e2e = float(ms_e2e_t)
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-entry-point device times: host-launched steps of the same process with CUDA events around every C-ABI call
    #      (kernels replayed from a CUDA graph cannot be bracketed), the last `nprof` of nprof + 3 ---------------------------
    nprof = 3
    if training and use_graph:
        t._cg = None
    for i in range(3):
        step(i)
    barrier()
    _lib.time_calls("all")
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for i in range(nprof):
        step(i)
    pe1.record()
    table_raw = _lib.timed_calls()
    prof_step_ms = pe0.elapsed_time(pe1) / nprof
    _lib.time_calls(None)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * B * steps / (ms / 1e3)
    e2e_value = world * B * e2e_steps / (ms_e2e / 1e3)
    step_ms = ms / steps
    pk = peaks()
    sustained = ms > 3000.0                                   # a timed region of seconds runs under the power cap
    rows = kernel_table(table_raw, nprof, prof_step_ms, pk, sustained)
    dom = next((r for r in rows if "bound" in r), None)
    h2d = sum(int(np.asarray(a).nbytes) for a in feed_batches[0] if a is not None)
    if training:
        d2h = 4
    else:
        T_ = 1 + (L_SAMPLES - m["window_size"]) // m["hop_size"]
        d2h = B * S * ((T_ - 1) * m["hop_size"] + m["window_size"]) * 4
    linear = (kind == "front_train" and precision == "bf16" and S == 2 and not args.stock_front and
              all(bool(np.array_equal(hb[0], hb[1][:, 0] + hb[1][:, 1])) for hb in host_batches))
    line = {
        "metric": METRIC if training else METRIC_INFER, "value": value,
        "unit": "mixtures/s", "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if precision == "fp32" else "bf16", "data": "synthetic",
        "config": {"workload": c["workload"], "baseline_config": cid, "batch_per_gpu": B, "global_batch": B * world,
                   "parallelism": f"dp{world}" if training else f"replicas x{world} (no collective)",
                   "host_numa_node": numa_node,
                   "precision": precision, "cuda_graph": use_graph,
                   "gradient_exchange": ("none (1 process)" if world == 1 or not training else
                                         "one flat all-reduce after backward" if args.flat_allreduce else
                                         "per-layer buckets all-reduced from the backward pass" +
                                         (", captured in the step graph" if use_graph else "")),
                   "inputs": ("sources + speaker ids from the host; the mixture is built on the device (amss_prepare_inputs)"
                              if training and not args.ship_mix else "mixture (+ sources) from the host"),
                   "timed_region_s": ms / 1e3,
                   "l2": "per-step working set (embeddings V + saved activations) exceeds the 126 MB L2; no explicit flush"},
        "e2e": {"value": e2e_value, "unit": "mixtures/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if kind == "front_train":
        line["config"]["front"] = ("mixture rows by linearity of the convolution (x_mix == x_0 + x_1 verified on the device "
                                   "per batch; stock kernel otherwise)" if linear else "stock: all B*(S+1) signals convolved")
    if dom is not None:
        rl = {"kernel": f"{dom['entry']}: {dom['what']}", "bound": dom["bound"], "achieved": dom["achieved"],
              "peak": dom["peak"], "unit": dom["unit"], "frac": dom["frac"], "traffic": None,
              "peak_source": pk["source"] + (" (sustained cuBLAS bf16)" if sustained and dom["bound"] == "tensor" else
                                             " (burst cuBLAS bf16)" if dom["bound"] == "tensor" else " (copy bandwidth)"),
              "ms_per_launch": dom["ms_per_step"] / max(1.0, dom["calls_per_step"]), "share_of_step": dom["share_of_step"],
              "work_per_step": dom["work_per_step"],
              "timing": f"CUDA events around the C-ABI call in {nprof} host-launched steps after the timed region "
                        f"({prof_step_ms:.3f} ms/step there)"}
        if dom["entry"] == "amss_filterbank_analysis_mix_fwd" and linear:
            # the library derives the mixture rows from the source rows (linearity): it EXECUTES S/(S+1) of 8(d)'s flops
            ex = S / (S + 1.0)
            rl.update(achieved_executed=dom["achieved"] * ex, frac_executed=dom["frac"] * ex,
                      executed_over_algorithmic_flops=ex,
                      note="achieved follows SURVEY 8(d) (2*L*W*N per signal, S+1 signals per mixture); the kernel executes "
                           "S of the S+1 convolutions, so frac_executed is the hardware utilisation; compute-bound "
                           "(ncu at 128 mixtures: tensor pipe active 95.9 % of the cycles, DRAM 66 MB read + 237 MB written of the "
                           "66 + 295 MB algorithmic: profiles/r02r_ncu_full_analysis_pair_B128.csv), DRAM traffic not a limiter")
        if "achieved_tfma" in dom:
            rl.update(limiter=dom["limiter"], achieved_tfma=dom["achieved_tfma"], smem_roof_tfma=dom["smem_roof_tfma"],
                      frac_of_smem_roof=dom["frac_of_smem_roof"],
                      note="a sparse CUDA-core kernel: its algorithmic HBM bytes (SURVEY 8(d)) are a few MB per step, so `frac` says "
                           "nothing about it; the multiply-adds are fed from shared memory and bound by that pipe "
                           "(achieved_tfma / smem_roof_tfma)")
        if dom["entry"] == "amss_kmeans_fit":
            a = table_raw["amss_kmeans_fit"][0][0]
            moved = float(a[3]) * (a[8] + 2) * a[4] * a[5] * 4.0 * len(table_raw["amss_kmeans_fit"]) / nprof
            phys = moved / (dom["ms_per_step"] / 1e3) / 1e9
            # `achieved` / `frac` are the bytes the kernel MOVES: (iters+2) passes over X, each labelling every try at once.
            # SURVEY 8(d) counts those passes PER TRY (the reference tiles X nb_tries times): that figure is kept beside it, it
            # is not a hardware fraction.  What bounds a pass is the fixed cost of its small tcgen05 MMAs (18 x N=32 split
            # products + 8 one-hot products per 128-point tile) and the SM's instruction issue, ~1700 clk per 20 KB tile
            # (profiles/r02l_kmeans_tile_profile.txt)
            rl.update(achieved=phys, frac=phys / dom["peak"], achieved_survey_per_try=dom["achieved"],
                      frac_survey_per_try=dom["frac"], limiter="tcgen05 issue (fixed cost per small MMA) + instruction issue",
                      note="achieved = bytes moved: (iters+2) passes over X, every pass labels all tries; SURVEY 8(d)'s figure "
                           "counts the passes per try (achieved_survey_per_try) and exceeds the HBM peak by construction")
        line["roofline"] = rl
    line["kernels"] = [{k: (round(v, 6) if isinstance(v, float) else v) for k, v in r.items() if k != "what"} for r in rows[:10]]
    # BLSTM-path tensor utilisation (SURVEY 7 / BASELINE metric "BLSTM TC util %"): algorithmic flops of the BLSTM stack +
    # head + their gradients / device time of those entry points / measured tensor peak
    bl = [r for r in rows if r["entry"] in BLSTM_PATH and "work_per_step" in r]
    if bl:
        fl, tm = sum(r["work_per_step"] for r in bl), sum(r["ms_per_step"] for r in bl)
        peak = pk["tf_sust"] if sustained else pk["tf_burst"]
        line["blstm_tc_util_pct"] = 100.0 * fl / (tm / 1e3) / 1e12 / peak
        line["blstm_path"] = {"flops_per_step": fl, "ms_per_step": tm, "tflops": fl / (tm / 1e3) / 1e12, "peak": peak,
                              "entries": [r["entry"] for r in bl]}
    line["final_result"] = results[-1] if results else None
    if world == 1 and not args.no_cpu:
        cb, cs, cw = CPU_SAMPLE[cid]
        try:
            v, per_step, threads = cpu_steps(cid, cb, cs, cw)
            line["cpu_baseline"] = {"value": v, "unit": "mixtures/s", "cores": threads, "kind": "port",
                                    "sample": f"{cs} steps of batch {cb} after {cw} warm-up; {CPU_NOTE}"}
        except Exception as ex:  # the baseline is reported, never required
            line["cpu_baseline"] = {"value": None, "unit": "mixtures/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {ex}"}
        if cid in (1, 2) and not args.no_parity:
            try:
                line["parity"] = parity_vs_oracle(cid, precision)
            except Exception as ex:
                line["parity"] = {"failed": str(ex)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, choices=sorted(CONFIGS), default=2, help="BASELINE.json config (1-based)")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--seconds", type=float, default=0.0, help="time at least this many seconds of steps (sustained run)")
    ap.add_argument("--batch", type=int, default=0, help="mixtures per GPU per step (default: the config's)")
    ap.add_argument("--precision", choices=["fp32", "bf16"], default="bf16",
                    help="bf16 = tcgen05 kernels; fp32 = the SIMT parity kernels")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity report of the cpu_baseline leg")
    ap.add_argument("--ship-mix", action="store_true",
                    help="ship the mixture from the host too (default: only the sources; the device builds x_mix = sum)")
    ap.add_argument("--no-cuda-graph", dest="cuda_graph", action="store_false",
                    help="launch every kernel of the step from the host instead of replaying forward + backward from a CUDA "
                         "graph (Trainer.enable_cuda_graph)")
    ap.add_argument("--flat-allreduce", action="store_true",
                    help="A/B (multi-GPU): one all-reduce of the whole gradient buffer after backward, outside the step graph, "
                         "instead of the per-layer buckets launched from the backward pass and captured into the graph")
    ap.add_argument("--stock-front", action="store_true",
                    help="A/B: run all B*(S+1) signals through the stock analysis kernel instead of deriving the mixture "
                         "rows from the source rows by linearity (sets AMSS_NO_LINEAR_MIX=1)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
