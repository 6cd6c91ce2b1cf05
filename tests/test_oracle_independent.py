"""Independent pins for the two restated librosa routines that round 1 checked only against themselves:
``power_to_db`` and ``spectral_contrast`` (oracle/librosa_shim.py).

* ``power_to_db`` against ``torchaudio.functional.amplitude_to_DB`` (an independent implementation of the same definition:
  10 log10(max(amin, x)) - 10 log10(max(amin, ref)), clamped at ``max - top_db``), incl. ``ref=np.max`` as manager.py:223 calls it.
* ``spectral_contrast`` against a brute-force float64 implementation written here from librosa's documentation
  (Jiang et al. 2002: per octave band -- [0, fmin], [fmin, 2 fmin], ... , the last one open to Nyquist -- the mean of the top
  ``quantile`` of the bins minus the mean of the bottom ``quantile``, both in dB), frame by frame with Python lists and sorts,
  sharing no code with the shim; the dB step goes through torchaudio.

Reference call sites: sygnals/core/features/manager.py:205-227, sygnals/core/features/frequency_domain.py:147-212.
"""
import math

import numpy as np
import pytest

from oracle import librosa_shim as shim

torch = pytest.importorskip("torch")
taF = pytest.importorskip("torchaudio.functional")


def _ta_power_to_db(x, ref_value, top_db=80.0, amin=1e-10):
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
    return taF.amplitude_to_DB(t, 10.0, amin, math.log10(max(amin, ref_value)), top_db).numpy()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_power_to_db_vs_torchaudio(seed):
    rng = np.random.default_rng(seed)
    S = rng.random((40, 101)) * 10.0 ** rng.integers(-14, 3, (40, 101))          # spans amin and the -80 dB clamp
    S[3, 7] = 0.0
    np.testing.assert_allclose(shim.power_to_db(S, ref=np.max), _ta_power_to_db(S, float(S.max())), rtol=0, atol=1e-10)
    np.testing.assert_allclose(shim.power_to_db(S), _ta_power_to_db(S, 1.0), rtol=0, atol=1e-10)
    np.testing.assert_allclose(shim.power_to_db(S, ref=np.max, top_db=None), _ta_power_to_db(S, float(S.max()), top_db=None), rtol=0, atol=1e-10)
    one_col = S[:, :1]                                                            # a single frame: its own maximum is the reference
    np.testing.assert_allclose(shim.power_to_db(one_col, ref=np.max), _ta_power_to_db(one_col, float(one_col.max())), rtol=0, atol=1e-10)


def brute_force_contrast(S, sr, n_fft, n_bands=6, fmin=200.0, quantile=0.02):
    """float64 [n_bands + 1, T]; S: magnitudes [1 + n_fft/2, T]."""
    B, T = S.shape
    bin_hz = [i * (sr / 2.0) / (B - 1) for i in range(B)]                          # rfft bin frequencies
    edges = [0.0] + [fmin * 2.0 ** i for i in range(n_bands + 1)]
    peaks = [[0.0] * T for _ in range(n_bands + 1)]
    valleys = [[0.0] * T for _ in range(n_bands + 1)]
    for k in range(n_bands + 1):
        lo, hi = edges[k], edges[k + 1]
        members = [i for i in range(B) if lo <= bin_hz[i] <= hi]
        if k > 0:
            members = [members[0] - 1] + members                                  # every band but the first also takes the bin below it
        if k == n_bands:
            members = members + list(range(members[-1] + 1, B))                   # the top band runs up to Nyquist
        n_keep = max(int(np.rint(quantile * len(members))), 1)                    # bins averaged at either end
        if k < n_bands:
            members = members[:-1]                                                # ... and all but the top band drop their last bin
        for t in range(T):
            col = sorted(float(S[i, t]) for i in members)
            valleys[k][t] = sum(col[:n_keep]) / len(col[:n_keep])
            peaks[k][t] = sum(col[-n_keep:]) / len(col[-n_keep:])
    P, V = np.array(peaks), np.array(valleys)
    return _ta_power_to_db(P, 1.0) - _ta_power_to_db(V, 1.0)                      # each array clamped 80 dB below its own maximum


@pytest.mark.parametrize("sr,n_fft,n_bands,fmin,q", [(44100, 2048, 6, 200.0, 0.02), (22050, 2048, 6, 200.0, 0.02), (16000, 512, 4, 100.0, 0.1),
                                                      (8000, 256, 3, 300.0, 0.5), (48000, 1024, 6, 200.0, 0.02)])
def test_spectral_contrast_vs_brute_force(sr, n_fft, n_bands, fmin, q):
    rng = np.random.default_rng(n_fft + n_bands)
    B, T = 1 + n_fft // 2, 9
    S = rng.random((B, T)) * 10.0 ** rng.uniform(-6, 1, (B, T))
    S[:, 3] = 0.0                                                                 # silent frame: amin floor on both sides
    S[10:20, 4] = S[10, 4]                                                        # ties inside a band
    S[:, 5] = np.round(S[:, 5], 1)                                                # many ties / zeros
    got = shim.feature.spectral_contrast(S=S, sr=sr, n_fft=n_fft, n_bands=n_bands, fmin=fmin, quantile=q)
    ref = brute_force_contrast(S, sr, n_fft, n_bands, fmin, q)
    assert got.shape == ref.shape == (n_bands + 1, T)
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-9)


def test_contrast_band_layout_matches_the_engine_plan():
    """The brute-force band membership equals the (first bin, count, quantile count) table the CUDA plan is built from."""
    from backends import get_engine
    lib = get_engine("emu").lib
    for sr, n_fft, nb, fmin, q in [(44100, 2048, 6, 200.0, 0.02), (22050, 2048, 6, 200.0, 0.02), (16000, 1024, 5, 150.0, 0.05)]:
        B = 1 + n_fft // 2
        bin_hz = [i * (sr / 2.0) / (B - 1) for i in range(B)]
        edges = [0.0] + [fmin * 2.0 ** i for i in range(nb + 1)]
        lo_, cnt_, nq_ = lib.debug_contrast_bands(sr, n_fft, nb, fmin, q)
        for k in range(nb + 1):
            m = [i for i in range(B) if edges[k] <= bin_hz[i] <= edges[k + 1]]
            if k > 0:
                m = [m[0] - 1] + m
            if k == nb:
                m = m + list(range(m[-1] + 1, B))
            n_keep = max(int(np.rint(q * len(m))), 1)
            if k < nb:
                m = m[:-1]
            assert (lo_[k], cnt_[k], nq_[k]) == (m[0], len(m), n_keep)
