"""STFT magnitude / power through the TMA-staged ring kernel (syg_stft_ring.cuh) against the oracle's compute_stft
(sygnals/core/dsp.py:167-229 on the librosa shim).  The tests assert that the ring kernel is what ran (syg_debug_last_stft_path),
cover every n_fft it serves, partial rounds, units shorter than a frame, hops that are not n_fft/4, center=False, batches that
straddle rounds, and the fall-back conditions (odd hop, reflect padding, complex output, unaligned units)."""
import numpy as np
import pytest

from backends import BACKENDS, get_engine
from oracle import sygnals_oracle as orc
from sygnals_b200 import _ffi
from sygnals_b200.utils import synth


@pytest.fixture(params=BACKENDS)
def eng(request):
    return get_engine(request.param)


def _aligned(a):
    """copy of `a` whose data pointer is 16-byte aligned (the ring kernel's eligibility rule for bulk copies)"""
    buf = np.empty(a.size + 4, dtype=a.dtype)
    off = (-buf.ctypes.data // a.itemsize) % (16 // a.itemsize)
    out = buf[off:off + a.size].reshape(a.shape)
    out[...] = a
    assert out.ctypes.data % 16 == 0
    return out


def _oracle_mag(y, n_fft, hop, center=True, power=False):
    S = np.abs(orc.compute_stft(y.astype(np.float64), n_fft=n_fft, hop_length=hop, win_length=n_fft, window="hann", center=center,
                                pad_mode="constant"))
    return S * S if power else S


def _check(got, ref, power):
    # power: |dP| <= 1e-4 P + 1e-6 max_bin P per frame; magnitude: the same bound on the amplitude scale
    if power:
        tol = 1e-4 * ref + 1e-6 * ref.max(axis=-2, keepdims=True)
    else:
        tol = 5e-5 * ref + 1e-6 * ref.max(axis=-2, keepdims=True) + 1e-12
    bad = np.abs(got - ref) > tol
    assert not bad.any(), f"{bad.sum()} bins out of tolerance, worst excess {(np.abs(got - ref) - tol).max():.3e}"


@pytest.mark.parametrize("n_fft,hop,L,n,kind", [
    (256, 64, 16000, 3, _ffi.OUT_MAGNITUDE),        # cfg2 shapes: T = 251 (rounds of 64: three full + one partial)
    (512, 128, 16000, 2, _ffi.OUT_MAGNITUDE),
    (1024, 256, 16000, 2, _ffi.OUT_MAGNITUDE),
    (2048, 512, 16000, 2, _ffi.OUT_POWER),
    (512, 160, 16000, 5, _ffi.OUT_POWER),           # speech-commands hop (not n_fft/4)
    (256, 64, 100, 7, _ffi.OUT_MAGNITUDE),          # units shorter than a frame: every frame touches both paddings
    (1024, 2, 1200, 2, _ffi.OUT_MAGNITUDE),         # tiny hop: many rounds per unit, stage offsets inside the left padding
    (2048, 700, 44100, 1, _ffi.OUT_MAGNITUDE),      # hop not a divisor of anything, one long unit
])
def test_ring_vs_oracle(eng, n_fft, hop, L, n, kind):
    y = np.stack([synth.mixture(L, 16000, seed=31 * n_fft + i) for i in range(n)])
    if n > 2:
        y[1] = synth.edge_clip("impulse", L, 16000)
        y[2] = 0.0
    y = _aligned(y.astype(np.float32))
    out = eng.stft_host(y.reshape(-1), eng.units_clips(n, L), n_fft, hop, n_fft, 0, True, 0, kind)
    assert eng.lib.dll.syg_debug_last_stft_path() == 1, "the ring kernel did not run"
    for i in range(n):
        _check(out[i].astype(np.float64), _oracle_mag(y[i], n_fft, hop, power=(kind == _ffi.OUT_POWER)), kind == _ffi.OUT_POWER)


def test_ring_center_false_and_overlapping_units(eng):
    """center=False (no padding at all) and analytic units that overlap (stride < unit_len, zero tail past total_len)."""
    n_fft, hop, L = 512, 128, 4096
    y = _aligned(synth.mixture(3 * 2048 + 1024, 22050, seed=5))
    u = eng.units_clips(4, L, total_len=y.size, stride=2048)                      # the last unit is cut by total_len -> zero tail
    out = eng.stft_host(y, u, n_fft, hop, n_fft, 0, False, 0, _ffi.OUT_MAGNITUDE)
    assert eng.lib.dll.syg_debug_last_stft_path() == 1
    for i in range(4):
        seg = np.zeros(L, dtype=np.float32)
        v = max(0, min(L, y.size - i * 2048))
        seg[:v] = y[i * 2048: i * 2048 + v]
        _check(out[i].astype(np.float64), _oracle_mag(seg, n_fft, hop, center=False), False)


def test_ring_equals_register_staged_kernel(eng):
    """Same spectra from the ring kernel and from the register-staged warp kernel (they share the FFT; the split differs in the
    last bits only), and the documented fall-backs take the other kernels."""
    n_fft, hop, L = 1024, 256, 8000
    y = _aligned(np.stack([synth.mixture(L, 16000, seed=70 + i) for i in range(3)]).astype(np.float32))
    a = eng.stft_host(y.reshape(-1), eng.units_clips(3, L), n_fft, hop, n_fft, 0, True, 0, _ffi.OUT_MAGNITUDE)
    assert eng.lib.dll.syg_debug_last_stft_path() == 1
    b = eng.stft_host(y.reshape(-1), eng.units_clips(3, L), n_fft, hop + 1, n_fft, 0, True, 0, _ffi.OUT_MAGNITUDE)   # odd hop
    assert eng.lib.dll.syg_debug_last_stft_path() == 2
    c = eng.stft_host(y.reshape(-1), eng.units_clips(3, L), n_fft, hop, n_fft, 0, True, 1, _ffi.OUT_MAGNITUDE)       # reflect padding
    assert eng.lib.dll.syg_debug_last_stft_path() == 2
    d = eng.stft_host(y.reshape(-1), eng.units_clips(3, L), n_fft, hop, n_fft, 0, True, 0, _ffi.OUT_COMPLEX)
    assert eng.lib.dll.syg_debug_last_stft_path() == 2
    np.testing.assert_allclose(a, np.abs(d), rtol=2e-5, atol=1e-6 * np.abs(d).max())
    interior = slice(2, a.shape[2] - 2)                                          # reflect and zero padding agree away from the edges
    np.testing.assert_allclose(a[:, :, interior], c[:, :, interior], rtol=2e-5, atol=1e-6 * a.max())
    assert b.shape[2] == 1 + L // (hop + 1)
    # a sample buffer that is not 16-byte aligned (device entry point; the host pipeline stages into aligned buffers)
    yo = np.zeros(y.size + 1, dtype=np.float32)
    if getattr(eng, "test_backend", None) == "emu":
        yo = _aligned(yo)
        yo[1:] = y.reshape(-1)
        T = 1 + L // hop
        out = np.zeros((3, n_fft // 2 + 1, T), dtype=np.float32)
        eng.stft_dev(yo.ctypes.data + 4, eng.units_clips(3, L), n_fft, hop, n_fft, 0, True, 0, _ffi.OUT_MAGNITUDE, out.ctypes.data)
        assert eng.lib.dll.syg_debug_last_stft_path() == 2
        np.testing.assert_allclose(out, a, rtol=2e-5, atol=1e-6 * a.max())


@pytest.mark.parametrize("n_fft,hop,L,n,kind,pad", [
    (4096, 1024, 16000, 2, _ffi.OUT_MAGNITUDE, 0),     # cfg2 shapes: T = 16 (two rounds of 8 frames per clip)
    (8192, 2048, 16000, 3, _ffi.OUT_MAGNITUDE, 0),     # T = 8; rounds of 4 frames straddle clips when n is odd
    (4096, 700, 9001, 3, _ffi.OUT_POWER, 0),           # ragged length, hop not a divisor, odd positions (unaligned pairs)
    (8192, 1000, 30000, 1, _ffi.OUT_POWER, 1),         # reflect padding goes through the predicated loads
    (4096, 1024, 3000, 2, _ffi.OUT_MAGNITUDE, 0),      # unit shorter than the frame
])
def test_big_transforms_vs_oracle(eng, n_fft, hop, L, n, kind, pad):
    """n_fft 4096 / 8192 through stft_big_kernel (1024-point sub-FFTs + recombination) against the oracle's compute_stft."""
    y = np.stack([synth.mixture(L, 16000, seed=17 * n_fft + i) for i in range(n)]).astype(np.float32)
    if n > 2:
        y[1] = synth.edge_clip("impulse", L, 16000)
    out = eng.stft_host(y.reshape(-1), eng.units_clips(n, L), n_fft, hop, n_fft, 0, True, pad, kind)
    assert eng.lib.dll.syg_debug_last_stft_path() == 4, "the sub-FFT kernel did not run"
    for i in range(n):
        S = np.abs(orc.compute_stft(y[i].astype(np.float64), n_fft=n_fft, hop_length=hop, win_length=n_fft, window="hann", center=True,
                                    pad_mode="reflect" if pad else "constant"))
        _check(out[i].astype(np.float64), S * S if kind == _ffi.OUT_POWER else S, kind == _ffi.OUT_POWER)
    # complex output of the same transform still comes from the CTA-cooperative kernel and agrees
    c = eng.stft_host(y.reshape(-1), eng.units_clips(n, L), n_fft, hop, n_fft, 0, True, pad, _ffi.OUT_COMPLEX)
    assert eng.lib.dll.syg_debug_last_stft_path() == 3
    ref = np.abs(c) ** 2 if kind == _ffi.OUT_POWER else np.abs(c)
    np.testing.assert_allclose(out, ref, rtol=3e-4 if kind == _ffi.OUT_POWER else 1.5e-4, atol=2e-6 * ref.max())
