"""Seeded synthetic PCM used by fixtures, parity tests and bench.py.

make_signal(kind, channels, frames, bits, rate, seed) -> int64 [frames, channels]

  music        0.30 FS sine at 220(c+1) Hz + 0.09 FS sine at 1333+37c Hz + N(0, 0.003 FS)
  silence_lsb  music with 25 % digital silence and 25 % +-2 LSB noise spliced in, so the Golomb
               zero-run (dynGet) and k==1 branches are exercised
  bench        music with 2 s silence + 2 s +-2 LSB noise in every 30 s (the section 8d recipe)
  lsb          +-2 LSB noise only (the quiet-passage regime of the entropy coder)
  white        full-scale uniform noise (encoders fall back to escape elements)
  loud         near-full-scale two-tone with hard clipping and bursts (large residuals, escape codes)
"""
import numpy as np


def _music(ch, n, bits, rate, rng, t0=0):
    fs = float(2 ** (bits - 1))
    t = (np.arange(n, dtype=np.float64) + t0) / rate
    out = np.empty((n, ch), dtype=np.float64)
    for c in range(ch):
        out[:, c] = (0.30 * fs * np.sin(2 * np.pi * 220.0 * (c + 1) * t)
                     + 0.09 * fs * np.sin(2 * np.pi * (1333.0 + 37.0 * c) * t)
                     + rng.normal(0.0, 0.003 * fs, n))
    return out


def make_signal(kind, ch, n, bits, rate, seed, t0=0):
    rng = np.random.default_rng(seed)
    fs = 2 ** (bits - 1)
    if kind == 'lsb':  # +-2 LSB noise only: the quiet-passage regime (one run-length code per sample)
        return rng.integers(-2, 3, size=(n, ch), dtype=np.int64)
    if kind == 'white':
        return rng.integers(-fs, fs, size=(n, ch), dtype=np.int64)
    x = _music(ch, n, bits, rate, rng, t0)
    if kind == 'loud':
        x *= 3.0
        burst = rng.random(n) < 0.002
        x[burst] += rng.normal(0.0, 0.8 * fs, (int(burst.sum()), ch))
    x = np.clip(np.round(x), -fs, fs - 1).astype(np.int64)
    if kind == 'silence_lsb':
        a, b, c = int(0.20 * n), int(0.45 * n), int(0.70 * n)
        x[a:b] = 0
        x[b:c] = rng.integers(-2, 3, size=(c - b, ch))
    elif kind == 'bench':
        period, seg = 30 * rate, 2 * rate
        pos = (np.arange(n, dtype=np.int64) + t0) % period
        sil = (pos >= 10 * rate) & (pos < 10 * rate + seg)
        lsb = (pos >= 20 * rate) & (pos < 20 * rate + seg)
        x[sil] = 0
        x[lsb] = rng.integers(-2, 3, size=(int(lsb.sum()), ch))
    return x
