"""
Known-answer tests restated from the reference's own test-suite and applied to the oracle:
  /root/reference/tests/test_segmentation.py:54-169  (exact contents / counts / zero padding)
  /root/reference/tests/test_features_time.py:120-131 (crest factor edge cases)
  /root/reference/tests/test_features_freq.py:78-101,157-184 (centroid / rolloff edge cases)
  /root/reference/tests/test_features_manager.py (frame counts, names, short signals, errors)
The same contract is applied to the CUDA engine in tests/test_gpu_contract.py.
"""
import numpy as np
import pytest
from numpy.testing import assert_array_equal

from oracle import sygnals_oracle as O


@pytest.fixture
def long_signal():
    sr = 1000
    rng = np.random.default_rng(0)
    return rng.standard_normal(int(5.3 * sr)), sr


def test_segment_no_overlap_no_pad(long_signal):
    y, sr = long_signal
    segs = O.segment_fixed_length(y, sr, 1.0, overlap_ratio=0.0, pad=False)
    assert len(segs) == len(y) // 1000
    for i, s in enumerate(segs):
        assert s.dtype == np.float64 and len(s) == 1000
        assert_array_equal(s, y[i * 1000:(i + 1) * 1000])


def test_segment_overlap_no_pad(long_signal):
    y, sr = long_signal
    segs = O.segment_fixed_length(y, sr, 1.0, overlap_ratio=0.5, pad=False)
    n, start = 0, 0
    while start + 1000 <= len(y):
        n += 1
        start += 500
    assert len(segs) == n
    for i, s in enumerate(segs):
        assert_array_equal(s, y[i * 500:i * 500 + 1000])


def test_segment_padding(long_signal):
    y, sr = long_signal
    segs = O.segment_fixed_length(y, sr, 1.0, overlap_ratio=0.25, pad=True)
    n, start = 0, 0
    while start < len(y):
        n += 1
        start += 750
    assert len(segs) == n
    last = segs[-1]
    s0 = (n - 1) * 750
    orig = len(y) - s0
    assert len(last) == 1000 and orig > 0
    assert_array_equal(last[:orig], y[s0:])
    assert_array_equal(last[orig:], np.zeros(1000 - orig))


def test_segment_min_length_and_short(long_signal):
    y, sr = long_signal
    assert len(O.segment_fixed_length(y, sr, 1.0, 0.0, True, 0.4)) == 5
    short = y[:500]
    assert len(O.segment_fixed_length(short, sr, 1.0, pad=False)) == 0
    p = O.segment_fixed_length(short, sr, 1.0, pad=True)
    assert len(p) == 1 and len(p[0]) == 1000
    assert_array_equal(p[0][:500], short)
    assert_array_equal(p[0][500:], np.zeros(500))
    assert len(O.segment_fixed_length(short, sr, 1.0, pad=True, min_segment_length_sec=0.6)) == 0


def test_segment_invalid(long_signal):
    y, sr = long_signal
    for kw in (dict(segment_length_sec=0), dict(segment_length_sec=1.0, overlap_ratio=1.0),
               dict(segment_length_sec=1.0, overlap_ratio=-0.1)):
        with pytest.raises(ValueError):
            O.segment_fixed_length(y, sr, **kw)
    with pytest.raises(ValueError):
        O.segment_fixed_length(np.zeros((2, 10)), sr, 1.0)


def test_crest_factor_known_answers():
    assert O.crest_factor(np.zeros(100)) == 0.0
    assert O.crest_factor(np.ones(100) * 5.0) == pytest.approx(1.0, abs=0)
    t = np.linspace(0, 1, 1000, endpoint=False)
    assert O.crest_factor(np.sin(2 * np.pi * 5 * t)) == pytest.approx(np.sqrt(2), rel=0.05)
    assert O.crest_factor(np.array([1.0, -1.0])) == 1.0
    assert O.crest_factor(np.array([])) == 0.0
    assert O.peak_amplitude(np.array([-1.2, 0.5, 1.0, -0.8])) == 1.2


def test_centroid_rolloff_known_answers():
    f = np.array([0, 100, 200, 300], dtype=float)
    assert O.spectral_centroid(np.array([0, 0.5, 1.0, 0.5]), f) == pytest.approx(200.0)
    assert O.spectral_centroid(np.zeros(4), f) == 0.0
    assert O.spectral_centroid(np.array([]), np.array([])) == 0.0
    f5 = np.array([0, 100, 200, 300, 400], dtype=float)
    m5 = np.array([0, 1, 1, 0, 0], dtype=float)
    assert O.spectral_rolloff(m5, f5, 0.85) == 200.0
    assert O.spectral_rolloff(m5, f5, 0.40) == 100.0
    assert O.spectral_rolloff(np.zeros(5), f5) == 400.0
    freqs = np.linspace(0, 11025, 1025)
    flat = np.ones(1025)
    assert O.spectral_rolloff(flat, freqs) == pytest.approx(0.85 * 11025, rel=0.15)
    rng = np.random.default_rng(42)
    mag = 10.0 * np.exp(-0.5 * ((np.arange(1025) - 300) / 3.0) ** 2) + 1e-4 * rng.random(1025)
    assert O.spectral_centroid(mag, freqs) == pytest.approx(freqs[300], rel=0.10)
    with pytest.raises(ValueError):
        O.spectral_rolloff(m5, f5, 1.5)


def test_manager_contract():
    sr = 22050
    t = np.arange(sr) / sr
    y = 0.5 * np.sin(2 * np.pi * 440 * t)
    r = O.extract_features(y, sr, ["rms_energy", "spectral_centroid", "mfcc", "spectral_contrast"], 1024, 256)
    T = 1 + len(y) // 256
    assert list(r)[:3] == ["time", "rms_energy", "spectral_centroid"]
    assert [k for k in r if k.startswith("mfcc_")] == [f"mfcc_{i}" for i in range(13)]
    assert [k for k in r if k.startswith("contrast")] == [f"contrast_band_{i}" for i in range(6)] + ["contrast_delta"]
    assert all(v.dtype == np.float64 and len(v) == T for v in r.values())
    assert r["time"][0] == pytest.approx(512 / sr)            # quirk: n_fft//2 offset (manager.py:168)
    assert np.mean(r["rms_energy"][4:-4]) == pytest.approx(0.5 / np.sqrt(2), abs=0.05)
    # short signals (tests/test_features_manager.py:183-220)
    r = O.extract_features(y[:512], sr, ["rms_energy", "spectral_centroid"], 1024, 256)
    assert len(r["time"]) == 3 and not any(np.isnan(v).any() for v in r.values())
    r = O.extract_features(y[:100], sr, ["rms_energy"], 1024, 512)
    assert len(r["time"]) == 1
    with pytest.raises(ValueError, match="Unknown feature"):
        O.extract_features(y, sr, ["nope"], 1024, 256)
    assert list(O.extract_features(y, sr, [], 1024, 256)) == ["time"]
    with pytest.raises(ValueError):
        O.extract_features(np.zeros((2, 100)), sr, ["rms_energy"])
    # 20-coefficient run contains the 13-coefficient run (tests/test_features_cepstral.py:117)
    a = O.extract_features(y, sr, ["mfcc"], 1024, 256, feature_params={"mfcc": {"n_mfcc": 20}})
    b = O.extract_features(y, sr, ["mfcc"], 1024, 256)
    for i in range(13):
        np.testing.assert_allclose(a[f"mfcc_{i}"], b[f"mfcc_{i}"], atol=1e-6)


def test_stft_contract():
    sr = 22050
    t = np.arange(sr) / sr
    y = np.sin(2 * np.pi * 1000 * t)
    D = O.compute_stft(y, n_fft=1024, hop_length=256)
    assert D.shape[0] == 513 and D.dtype == np.complex128 and D.shape[1] == 1 + len(y) // 256
    k = np.argmax(np.abs(D).mean(axis=1))
    assert abs(k * sr / 1024 - 1000) < sr / 1024
    f, p = O.compute_psd_welch(y, fs=sr, nperseg=4096)
    assert abs(f[np.argmax(p)] - 1000) < sr / 4096
    with pytest.raises(ValueError):
        O.compute_stft(np.zeros((2, 4096)))
