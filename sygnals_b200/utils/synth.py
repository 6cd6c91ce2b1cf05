"""
Synthetic audio used by the parity tests and by ``bench.py``:
per unit a float32 mixture ``a * (0.4 sin(2 pi f1 t) + 0.4 logchirp(f0 -> f2) + 0.05 N(0,1))`` with
frequencies log-uniform in [50, 0.45 sr] and amplitude ``a`` log-uniform in [0.01, 1] (exercises the
-80 dB clamp), plus the mandated edge clips.  Everything is quantised to float32 *before* either the
oracle (which up-casts, like ``sygnals/core/audio/io.py:94-95``) or the engine sees it.
"""
from __future__ import annotations

import numpy as np

EDGE_KINDS = ("zeros", "dc", "impulse", "square", "tiny", "loud_tone")


def mixture(n: int, sr: int, seed: int) -> np.ndarray:
    """One deterministic float32 clip of ``n`` samples."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / sr
    lo, hi = np.log(50.0), np.log(0.45 * sr)
    f1, f0, f2 = np.exp(rng.uniform(lo, hi, 3))
    amp = np.exp(rng.uniform(np.log(0.01), 0.0))
    dur = max(n / sr, 1e-3)
    k = (f2 / f0) ** (1.0 / dur)
    phase = 2 * np.pi * f0 * ((k ** t - 1.0) / np.log(k)) if abs(k - 1.0) > 1e-12 else 2 * np.pi * f0 * t
    y = amp * (0.4 * np.sin(2 * np.pi * f1 * t) + 0.4 * np.sin(phase) + 0.05 * rng.standard_normal(n))
    return y.astype(np.float32)


def edge_clip(kind: str, n: int, sr: int) -> np.ndarray:
    y = np.zeros(n, dtype=np.float32)
    if kind == "zeros":
        pass
    elif kind == "dc":
        y[:] = 0.25
    elif kind == "impulse":
        y[n // 3] = 1.0
    elif kind == "square":
        period = max(2, sr // 220)
        y[:] = np.where((np.arange(n) // (period // 2)) % 2 == 0, 1.0, -1.0)
    elif kind == "tiny":
        y[:] = (1e-7 * np.sin(2 * np.pi * 1000.0 * np.arange(n) / sr)).astype(np.float32)
    elif kind == "loud_tone":
        y[:] = np.sin(2 * np.pi * 997.0 * np.arange(n) / sr).astype(np.float32)
    else:
        raise ValueError(kind)
    return y


def clip_batch(n_clips: int, n: int, sr: int, seed: int = 1234, edges: bool = True) -> np.ndarray:
    """``float32[n_clips, n]``; the first ``len(EDGE_KINDS)`` clips are the edge cases when ``edges``."""
    out = np.empty((n_clips, n), dtype=np.float32)
    for i in range(n_clips):
        if edges and i < len(EDGE_KINDS):
            out[i] = edge_clip(EDGE_KINDS[i], n, sr)
        else:
            out[i] = mixture(n, sr, seed + 7919 * i)
    return out


def long_signal(n: int, sr: int, seed: int = 4321, block_sec: float = 1.0) -> np.ndarray:
    """A long float32 recording made of per-block mixtures (different amplitude / pitch per block)
    with a silent block and a DC block inserted so segments straddle them."""
    blk = max(1, int(block_sec * sr))
    out = np.empty(n, dtype=np.float32)
    for b, s in enumerate(range(0, n, blk)):
        e = min(n, s + blk)
        if b % 11 == 5:
            out[s:e] = 0.0
        elif b % 11 == 9:
            out[s:e] = 0.1
        else:
            out[s:e] = mixture(e - s, sr, seed + 104729 * b)
    return out


def torch_mixture_(out, sr: int, seed: int = 1234, unit: int = 0):
    """Fill a CUDA float32 tensor ``out[n_units, n]`` (or 1-D) in place with the same *family* of
    signals (not bit-identical to :func:`mixture`; the bench copies its CPU sample from the device)."""
    import torch

    flat = out.view(-1) if out.dim() == 1 else out
    if flat.dim() == 1:
        n_units = max(1, flat.numel() // max(1, unit or sr))
        n = flat.numel() // n_units
        body = flat[: n_units * n].view(n_units, n)
    else:
        body = flat
        n_units, n = body.shape
    g = torch.Generator(device=out.device)
    g.manual_seed(seed)
    dev = out.device
    lo, hi = float(np.log(50.0)), float(np.log(0.45 * sr))
    chunk = max(1, (64 << 20) // max(1, n))
    t = torch.arange(n, device=dev, dtype=torch.float32) / sr
    for s in range(0, n_units, chunk):
        e = min(n_units, s + chunk)
        m = e - s
        f = torch.exp(torch.rand(m, 3, device=dev, generator=g) * (hi - lo) + lo)
        amp = torch.exp(torch.rand(m, 1, device=dev, generator=g) * float(np.log(100.0)) - float(np.log(100.0)))
        ph1 = (2 * np.pi) * torch.remainder(f[:, 0:1] * t[None, :], 1.0)
        dur = n / sr
        k = (f[:, 2:3] / f[:, 1:2]) ** (1.0 / dur)
        lk = torch.log(k)
        cyc = f[:, 1:2].double() * ((torch.exp(lk.double() * t[None, :].double()) - 1.0) / lk.double())
        ph2 = (2 * np.pi) * torch.remainder(cyc, 1.0).float()
        blk = amp * (0.4 * torch.sin(ph1) + 0.4 * torch.sin(ph2)
                     + 0.05 * torch.randn(m, n, device=dev, generator=g))
        body[s:e] = blk
    if flat.dim() == 1 and n_units * n < flat.numel():
        flat[n_units * n:] = 0.0
    return out


def torch_recording_(out, begin: int, sr: int, seed: int = 1234, chunk_blocks: int = 32):
    """Fill the 1-D CUDA float32 tensor ``out`` with samples ``[begin, begin + len(out))`` of ONE deterministic, unbounded synthetic
    recording: one-second blocks of the mixture family (sine + log-chirp + noise, amplitudes over 40 dB).  Position addressable:
    any slice taken by any rank equals the same slice of the whole recording (blocks are generated in fixed global chunks of
    ``chunk_blocks`` seconds, each from its own seeds), which is what the strong-scaling bench and the sharded tests need."""
    import torch

    n = int(out.numel())
    dev = out.device
    lo, hi = float(np.log(50.0)), float(np.log(0.45 * sr))
    t = torch.arange(sr, device=dev, dtype=torch.float32) / sr
    t64 = t.double()
    span = chunk_blocks * sr
    c0, c1 = begin // span, (begin + n + span - 1) // span
    for c in range(c0, c1):
        rng = np.random.default_rng([seed, c])
        f = torch.from_numpy(np.exp(rng.uniform(lo, hi, (chunk_blocks, 3))).astype(np.float32)).to(dev)
        amp = torch.from_numpy(np.exp(rng.uniform(np.log(0.01), 0.0, (chunk_blocks, 1))).astype(np.float32)).to(dev)
        g = torch.Generator(device=dev)
        g.manual_seed(seed * 1000003 + c)
        ph1 = (2 * np.pi) * torch.remainder(f[:, 0:1] * t[None, :], 1.0)
        lk = torch.log(f[:, 2:3] / f[:, 1:2]).double()                         # chirp over the one-second block
        cyc = f[:, 1:2].double() * ((torch.exp(lk * t64[None, :]) - 1.0) / lk)
        ph2 = (2 * np.pi) * torch.remainder(cyc, 1.0).float()
        blk = amp * (0.4 * torch.sin(ph1) + 0.4 * torch.sin(ph2) + 0.05 * torch.randn(chunk_blocks, sr, device=dev, generator=g))
        flat = blk.reshape(-1)
        g0 = c * span                                                          # global position of the chunk
        s, e = max(begin, g0), min(begin + n, g0 + span)
        out[s - begin:e - begin] = flat[s - g0:e - g0]
    return out
