"""Transform lengths that are NOT powers of two (syg_mixed.cuh: shared-memory Stockham FFT with radices 4/2/3/5/7/11/13) through the
C ABI: golden vectors of the UNMODIFIED reference (tests/golden/make_golden_mixed.py) and the oracle on seeded inputs.  Reference:
sygnals/core/dsp.py:167-229 (STFT), :434-560 (periodogram / Welch), sygnals/core/features/manager.py:78-445 (the reference's own
tests frame with frame_length=1000).  Same tolerances as tests/test_parity_cabi.py.  Each case runs on the CPU fiber emulator
(test infrastructure) here and on the B200 under the ``gpu`` marker."""
import numpy as np
import pytest

import cases
from backends import BACKENDS, get_engine
from oracle import sygnals_oracle as orc
from sygnals_b200 import _ffi
from sygnals_b200.utils import synth
from test_parity_cabi import check_rows, oracle_rows, power_close


@pytest.fixture(params=BACKENDS)
def eng(request):
    return get_engine(request.param)


def gold():
    return cases.load("mixed_lengths.npz")


def test_oracle_matches_golden_mixed_lengths():
    """CPU: the oracle reproduces the reference's outputs for these lengths bit for bit (no engine involved)."""
    g = gold()
    clip = synth.mixture(8000, 16000, seed=202)
    np.testing.assert_allclose(cases.checksum(clip), g["stft_in_checksum"], rtol=1e-12)
    D = orc.compute_stft(clip.astype(np.float64), n_fft=400, hop_length=160)
    assert np.array_equal(D.astype(np.complex64), g["D_400_160"])
    x = synth.long_signal(25600, 25600, seed=808, block_sec=0.5)
    np.testing.assert_allclose(cases.checksum(x), g["psd_in_checksum"], rtol=1e-12)
    f, p = orc.compute_psd_periodogram(x.astype(np.float64), fs=25600, window="hann")
    assert np.array_equal(p, g["periodogram_25600"])
    y = synth.long_signal(44100, 44100, seed=707)
    names, rows = oracle_rows(y, 44100, cases.CFG4_FEATURES, 1000, 250)
    assert names == [str(n) for n in g["env1000_names"]]
    assert np.array_equal(rows, g["env1000_rows"])


def test_golden_stft_mixed_lengths(eng):
    g = gold()
    clip = synth.mixture(8000, 16000, seed=202)
    u = eng.units_clips(1, len(clip))
    for key, kw in [("D_400_160", dict(n_fft=400, hop=160, win=400)),
                    ("D_1000_250_reflect", dict(n_fft=1000, hop=250, win=1000, pad_mode=_ffi.PAD_IDS["reflect"])),
                    ("D_441_147_nocenter", dict(n_fft=441, hop=147, win=441, center=False)),
                    ("D_1200_win900_300", dict(n_fft=1200, hop=300, win=900))]:
        extra = {k: v for k, v in kw.items() if k in ("pad_mode", "center")}
        D = eng.stft_host(clip, u, kw["n_fft"], kw["hop"], kw["win"], **extra)
        ref = g[key].astype(np.complex128)
        assert D[0].shape == ref.shape, key
        scale = np.abs(ref).max(axis=0, keepdims=True)
        assert (np.abs(D[0] - ref) <= 2e-6 * scale + 1e-12).all(), key
        pw = eng.stft_host(clip, u, kw["n_fft"], kw["hop"], kw["win"], out_kind=_ffi.OUT_POWER, **extra)
        mag = eng.stft_host(clip, u, kw["n_fft"], kw["hop"], kw["win"], out_kind=_ffi.OUT_MAGNITUDE, **extra)
        power_close(pw[0].astype(np.float64), np.abs(ref) ** 2)
        power_close(mag[0].astype(np.float64) ** 2, np.abs(ref) ** 2)
    if eng.test_backend == "gpu":
        assert eng.lib.dll.syg_debug_last_stft_path() == 5              # the mixed-radix kernel ran


def test_golden_speech_framing_400(eng):
    """25 ms / 10 ms framing at 16 kHz (frame_length 400 = 2^4 5^2): MFCC13 over 40 mels + RMS, edge clips included."""
    g = gold()
    clips = synth.clip_batch(8, 16000, 16000, seed=606, edges=True)
    np.testing.assert_allclose(cases.checksum(clips), g["speech400_in_checksum"], rtol=1e-12)
    p = _ffi.make_params(eng.lib, 16000, ["mfcc", "rms_energy"], 400, 160, feature_params={"mfcc": {"n_mels": 40}})
    out = eng.features_host(clips.ravel(), eng.units_clips(clips.shape[0], clips.shape[1]), p)
    ref = g["speech400_rows"]
    assert out.shape == ref.shape == (8, 14, 101)
    check_rows([str(n) for n in g["speech400_names"]], out, ref)


def test_golden_env_features_1000(eng):
    g = gold()
    sr = 44100
    y = synth.long_signal(sr, sr, seed=707)
    np.testing.assert_allclose(cases.checksum(y), g["env1000_in_checksum"], rtol=1e-12)
    p = _ffi.make_params(eng.lib, sr, cases.CFG4_FEATURES, 1000, 250)
    out = eng.features_host(y, eng.units_clips(1, len(y)), p)
    ref = g["env1000_rows"]
    assert out[0].shape == ref.shape
    check_rows([str(n) for n in g["env1000_names"]], out[0], ref, bin_hz=sr / 1000, nyq=sr / 2)


def test_golden_psd_second_of_25600(eng):
    """BASELINE config 5's unit as ONE transform: the periodogram of a 25 600-sample second (2^10 5^2), plus Welch with 1000-sample
    sub-segments and an odd nperseg (945 = 3^3 5 7) zero-padded to 1890 with 'spectrum' scaling."""
    g = gold()
    sr = 25600
    x = synth.long_signal(sr, sr, seed=808, block_sec=0.5)
    u = eng.units_clips(1, len(x))
    for key, args in [("periodogram_25600", (0, 25600, 0, 25600, True, 0)), ("welch_1000_500", (0, 1000, 500, 1000, True, 0)),
                      ("welch_945_100_1890_spectrum", (1, 945, 100, 1890, True, 1))]:
        psd = eng.psd_welch_host(x, u, sr, *args)
        ref = g[key]
        assert psd[0].shape == ref.shape, key
        assert (np.abs(psd[0] - ref) <= 1e-4 * ref + 2e-6 * ref.max()).all(), key


ALL16 = ["mfcc", "spectral_contrast", "spectral_centroid", "spectral_rolloff", "rms_energy", "crest_factor", "peak_amplitude",
         "spectral_bandwidth", "spectral_flatness", "dominant_frequency", "mean_amplitude", "std_dev_amplitude", "zero_crossing_rate",
         "skewness", "kurtosis", "signal_entropy"]


@pytest.mark.parametrize("fl,hop,sr", [(400, 160, 16000), (1000, 250, 22050), (441, 110, 16000), (1200, 512, 44100), (2002, 500, 22050)])
def test_all_features_vs_oracle_mixed_lengths(eng, fl, hop, sr):
    """Every kernelled feature at even, odd and 7 * 11 * 13 frame lengths, several units per call (one of them silent)."""
    n = 3 * fl + 777
    y = np.stack([synth.long_signal(n, sr, seed=fl + c).astype(np.float32) for c in range(3)])
    y[1] = 0.0
    fp = {"mfcc": {"n_mels": 40, "n_mfcc": 13}}
    p = _ffi.make_params(eng.lib, sr, ALL16, fl, hop, feature_params=fp)
    out = eng.features_host(y.ravel(), eng.units_clips(3, n), p)
    for c in range(3):
        names, ref = oracle_rows(y[c], sr, ALL16, fl, hop, feature_params=fp)
        assert out[c].shape == ref.shape
        check_rows(names, out[c], ref, bin_hz=sr / fl, nyq=sr / 2)


def test_unsupported_lengths_are_refused(eng):
    y = np.zeros(4000, dtype=np.float32)
    with pytest.raises(NotImplementedError, match="prime factors"):
        eng.stft_host(y, eng.units_clips(1, len(y)), 2 * 37, 37, 2 * 37)
    with pytest.raises(NotImplementedError, match="shared memory"):
        eng.psd_welch_host(np.zeros(40000, dtype=np.float32), eng.units_clips(1, 40000), 1.0, 0, 40000, 0, 40000, True, 0)


def test_pow2_above_8192_takes_the_same_route(eng):
    rng = np.random.default_rng(5)
    y = rng.standard_normal(40000).astype(np.float32)
    pw = eng.stft_host(y, eng.units_clips(1, len(y)), 16384, 4096, 16384, out_kind=_ffi.OUT_POWER)
    ref = np.abs(orc.compute_stft(y.astype(np.float64), n_fft=16384, hop_length=4096)) ** 2
    power_close(pw[0].astype(np.float64), ref)


@pytest.mark.gpu
def test_mirrors_serve_mixed_lengths_on_gpu():
    """The reference-named entry points (sygnals_b200.core.*) with lengths the reference's own tests use: compute_stft(n_fft=400),
    extract_features(frame_length=1000) -> DataFrame, compute_psd_periodogram of a 25 600-sample second; the plugin wrapper keeps
    them on the engine (its reference fallback must not be reached)."""
    from sygnals_b200 import plugin as plg
    from sygnals_b200.core import dsp
    sr = 16000
    y = synth.mixture(8000, sr, seed=31).astype(np.float64)
    D = dsp.compute_stft(y, n_fft=400, hop_length=160)
    ref = orc.compute_stft(y, n_fft=400, hop_length=160)
    assert D.dtype == np.complex128 and D.shape == ref.shape
    assert np.abs(D - ref).max() <= 2e-6 * np.abs(ref).max()
    p = plg.SygnalsB200Plugin()
    fn = p.make_extract_features(original=lambda *a, **k: pytest.fail("must not reach the reference"))
    df = fn(y, sr, ["mfcc", "rms_energy", "spectral_centroid"], frame_length=1000, hop_length=250)
    r = orc.extract_features(y, sr, ["mfcc", "rms_energy", "spectral_centroid"], frame_length=1000, hop_length=250)
    assert df.index.name == "time" and list(df.columns) == [k for k in r if k != "time"]
    for k in df.columns:
        tol = dict(atol=1e-3, rtol=0) if k.startswith("mfcc") else dict(rtol=1e-5, atol=1e-6 * sr / 2)
        np.testing.assert_allclose(df[k].to_numpy(), r[k], **tol)
    x = synth.long_signal(25600, 25600, seed=808, block_sec=0.5).astype(np.float64)
    f, pxx = dsp.compute_psd_periodogram(x, fs=25600.0)
    fr, pr = orc.compute_psd_periodogram(x, fs=25600.0)
    np.testing.assert_array_equal(f, fr)
    assert (np.abs(pxx - pr) <= 1e-4 * pr + 2e-6 * pr.max()).all()
