"""Synthetic "LibriSpeech-shaped" mixtures: the data contract of the
reference's pipeline -- (mix [B,L], non_mix [B,S,L], ind [B,S]) with mix = sum of the sources
(data/dataset.py:462-468) and distinct speaker ids inside a mixture (data/dataset.py:473-480) --
without its TFRecord reader.  Host-side numpy; the trainer moves batches to the GPU from pinned
memory."""
import numpy as np


def synthetic_mixtures(B, S, L, seed=42, fs=16000, tot_speakers=251):
    """Speech-like sources: 12 harmonics of a slowly varying f0 in [90, 250] Hz with 1/k roll-off,
    low-passed noise, a 3-6 Hz syllabic envelope with ~25 % silence, RMS 0.05."""
    rng = np.random.RandomState(seed)
    t = np.arange(L, dtype=np.float64) / float(fs)
    R = B * S
    f0 = rng.uniform(90, 250, size=(R, 1))
    vib = 1.0 + 0.05 * np.sin(2 * np.pi * rng.uniform(0.5, 2.0, size=(R, 1)) * t + rng.uniform(0, 6.28, size=(R, 1)))
    phase = 2 * np.pi * np.cumsum(f0 * vib, axis=1) / fs
    sig = np.zeros((R, L))
    for k in range(1, 13):
        sig += np.sin(k * phase + rng.uniform(0, 6.28, size=(R, 1))) / k
    noise = rng.randn(R, L)
    c = np.cumsum(np.pad(noise, ((0, 0), (8, 0))), axis=1)
    sig += 0.3 * (c[:, 8:] - c[:, :-8]) / 8.0                    # 8-tap moving average
    env = 0.5 * (1 + np.sin(2 * np.pi * rng.uniform(3, 6, size=(R, 1)) * t + rng.uniform(0, 6.28, size=(R, 1))))
    sig *= np.clip((env - 0.25) / 0.75, 0.0, 1.0)
    sig *= 0.05 / (np.sqrt(np.mean(sig ** 2, axis=1, keepdims=True)) + 1e-12)
    non_mix = sig.reshape(B, S, L).astype(np.float32)
    ind = np.stack([rng.choice(tot_speakers, size=S, replace=False) for _ in range(B)]).astype(np.int32)
    return non_mix.sum(1).astype(np.float32), non_mix, ind


class SyntheticStream:
    """Endless stream of (mix, non_mix, ind) host batches; seed = base + 1000*rank + step."""

    def __init__(self, B, S, L, seed=42, rank=0, pool=4):
        self.B, self.S, self.L, self.seed, self.rank = B, S, L, seed, rank
        # generating speech-like audio costs ~10 ms per source on the host: keep a small pool of
        # distinct batches and cycle through it (the kernels' work does not depend on the values)
        self.pool = [synthetic_mixtures(B, S, L, seed + 1000 * rank + i) for i in range(pool)]
        self.i = 0

    def __iter__(self):
        return self

    def __next__(self):
        b = self.pool[self.i % len(self.pool)]
        self.i += 1
        return b
