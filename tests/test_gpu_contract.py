"""
The reference's own test-suite for the hot path, restated against the ENGINE through its drop-in mirror
(``sygnals_b200.core.*``: the same names and signatures the reference's tests import), once per build of the kernels
(``emu`` here, ``gpu`` on the B200 box).  Same assertions as tests/test_oracle_known_answers.py applies to the oracle:

  /root/reference/tests/test_features_manager.py:53-235   (types, dtypes, names, frame counts, short signals, errors)
  /root/reference/tests/test_features_cepstral.py:94-117  (n_mfcc prefix property)
  /root/reference/tests/test_dsp.py:79-91                 (compute_stft shape / dtype / energy bin)
  /root/reference/tests/test_audio_features.py:108-116    (rms_energy of a sine / of silence, +-0.05)
  /root/reference/tests/test_segmentation.py:54-169       (exact contents / counts / zero padding)

Features the reference computes with sequential CPU algorithms (hnr / jitter / shimmer, hence ``'all'``) have no kernel: the
mirror refuses them with NotImplementedError and the plugin routes those calls to the reference (tests/test_plugin.py).
"""
import numpy as np
import pandas as pd
import pytest
from numpy.testing import assert_allclose, assert_array_equal
from scipy.signal import chirp

from backends import BACKENDS, get_engine
from sygnals_b200 import _ffi
from sygnals_b200.core import dsp, segmentation
from sygnals_b200.core.features import manager
from sygnals_b200.core.features.manager import extract_features


@pytest.fixture(params=BACKENDS)
def engine(request):
    """Route the mirror's process-wide engine to the requested build for the duration of one test."""
    eng = get_engine(request.param)
    saved = dict(_ffi._engines)
    _ffi._engines.clear()
    _ffi._engines[0] = eng
    eng.test_backend = request.param
    yield eng
    _ffi._engines.clear()
    _ffi._engines.update(saved)


@pytest.fixture
def sample_audio_long():
    """tests/test_features_manager.py:21-31 (librosa.chirp == scipy chirp, logarithmic, phi = -90 deg)."""
    sr, duration = 22050, 2.5
    t = np.linspace(0, duration, int(sr * duration), endpoint=False)
    tc = np.arange(int(np.ceil(duration * sr))) / sr
    sig = 0.4 * np.sin(2 * np.pi * 220.0 * t) + 0.4 * chirp(tc, 400, duration, 1000, method="logarithmic", phi=-90)
    return sig.astype(np.float64), sr


def expected_frames(n, hop, fl, center=True):
    if center:
        return 1 + n // hop
    return 0 if n < fl else 1 + (n - fl) // hop


# ------------------------------------------------------------------------------------------------ features manager
def test_extract_single_time_feature(engine, sample_audio_long):
    y, sr = sample_audio_long
    df = extract_features(y, sr, ["mean_amplitude"], frame_length=1024, hop_length=512, output_format="dataframe")
    assert isinstance(df, pd.DataFrame) and "mean_amplitude" in df.columns and df.index.name == "time"
    assert df["mean_amplitude"].dtype == np.float64
    assert len(df) == expected_frames(len(y), 512, 1024)


def test_extract_single_freq_feature_dict(engine, sample_audio_long):
    y, sr = sample_audio_long
    r = extract_features(y, sr, ["spectral_centroid"], frame_length=2048, hop_length=512, output_format="dict_of_arrays")
    assert isinstance(r, dict) and "time" in r and r["spectral_centroid"].dtype == np.float64
    T = expected_frames(len(y), 512, 2048)
    assert len(r["spectral_centroid"]) == T and len(r["time"]) == T


def test_extract_mfcc(engine, sample_audio_long):
    y, sr = sample_audio_long
    df = extract_features(y, sr, ["mfcc"], frame_length=2048, hop_length=512, feature_params={"mfcc": {"n_mfcc": 5}})
    cols = [f"mfcc_{i}" for i in range(5)]
    assert list(df.columns) == cols and all(df[c].dtype == np.float64 for c in cols)
    assert len(df) == expected_frames(len(y), 512, 2048)


def test_extract_spectral_contrast_dict(engine, sample_audio_long):
    y, sr = sample_audio_long
    r = extract_features(y, sr, ["spectral_contrast"], frame_length=2048, hop_length=512,
                         feature_params={"spectral_contrast": {"n_bands": 6}}, output_format="dict_of_arrays")
    keys = [f"contrast_band_{i}" for i in range(6)] + ["contrast_delta"]
    assert [k for k in r if k != "time"] == keys and all(r[k].dtype == np.float64 for k in keys)
    T = expected_frames(len(y), 512, 2048)
    assert len(r["time"]) == T and len(r["contrast_band_0"]) == T


def test_extract_multiple_spectrum_features_one_pass(engine, sample_audio_long):
    """The reference checks that the STFT is computed once for several spectrum features (test_features_manager.py:137-160);
    here they are ONE launch of the fused kernel: the frame kernel runs exactly once per call."""
    y, sr = sample_audio_long
    engine.profile_enable(True)
    engine.profile_read(reset=True)
    try:
        df = extract_features(y, sr, ["spectral_centroid", "spectral_flatness", "spectral_contrast"], frame_length=2048,
                              hop_length=1024)
        prof = engine.profile_read(reset=True)
    finally:
        engine.profile_enable(False)
    if engine.test_backend == "gpu":                    # launch accounting uses CUDA events: not present in the emulator build
        assert prof["frame"][1] == 1
    assert {"spectral_centroid", "spectral_flatness", "contrast_band_0"} <= set(df.columns)
    assert len(df) == expected_frames(len(y), 1024, 2048)


def test_extract_feature_with_params(engine, sample_audio_long):
    y, sr = sample_audio_long
    r85 = extract_features(y, sr, ["spectral_rolloff"], frame_length=1024, hop_length=512, output_format="dict_of_arrays")
    r95 = extract_features(y, sr, ["spectral_rolloff"], frame_length=1024, hop_length=512,
                           feature_params={"spectral_rolloff": {"roll_percent": 0.95}}, output_format="dict_of_arrays")
    assert np.nanmean(r95["spectral_rolloff"]) > np.nanmean(r85["spectral_rolloff"])


def test_extract_short_signal(engine, sample_audio_long):
    y, sr = sample_audio_long
    df = extract_features(y[:512].copy(), sr, ["rms_energy", "spectral_centroid"], frame_length=1024, hop_length=256)
    assert len(df) == 3 and list(df.columns) == ["rms_energy", "spectral_centroid"]
    assert not df.isnull().values.any()


def test_extract_very_short_signal(engine):
    df = extract_features(np.zeros(100, dtype=np.float64), 22050, ["rms_energy"], frame_length=1024, hop_length=512)
    assert isinstance(df, pd.DataFrame) and not df.empty and len(df) == 1


def test_extract_unknown_feature(engine, sample_audio_long):
    y, sr = sample_audio_long
    with pytest.raises(ValueError, match="Unknown feature\\(s\\) requested: \\['this_is_not_a_feature'\\]"):
        extract_features(y, sr, ["rms_energy", "this_is_not_a_feature"])


def test_extract_no_features(engine, sample_audio_long):
    y, sr = sample_audio_long
    df = extract_features(y, sr, [], output_format="dataframe")
    r = extract_features(y, sr, [], output_format="dict_of_arrays")
    assert isinstance(df, pd.DataFrame) and df.empty
    assert list(r) == ["time"]


def test_extract_every_engine_feature(engine, sample_audio_long):
    """'all' minus the three features without a kernel (test_features_manager.py:237-275 restated for the engine's set)."""
    y, sr = sample_audio_long
    feats = sorted(manager.ENGINE_FEATURES)
    df = extract_features(y, sr, feats, frame_length=1024, hop_length=512)
    assert {"mean_amplitude", "spectral_centroid", "mfcc_0", "contrast_band_0", "rms_energy"} <= set(df.columns)
    assert len(df) == expected_frames(len(y), 512, 1024)
    assert np.isfinite(df.values).all()
    for name in (["all"], ["hnr", "jitter", "shimmer"]):
        with pytest.raises(NotImplementedError, match="no CUDA kernel"):
            extract_features(y, sr, name, frame_length=1024, hop_length=512)


def test_mfcc_parameters_prefix_property(engine, sample_audio_long):
    """tests/test_features_cepstral.py:94-117: a 20-coefficient run contains the 13-coefficient run."""
    y, sr = sample_audio_long
    a = extract_features(y, sr, ["mfcc"], 1024, 256, feature_params={"mfcc": {"n_mfcc": 20}}, output_format="dict_of_arrays")
    b = extract_features(y, sr, ["mfcc"], 1024, 256, output_format="dict_of_arrays")
    assert len([k for k in a if k.startswith("mfcc_")]) == 20 and len([k for k in b if k.startswith("mfcc_")]) == 13
    for i in range(13):
        assert_allclose(a[f"mfcc_{i}"], b[f"mfcc_{i}"], atol=1e-6)


def test_rms_energy_sine_and_silence(engine):
    """tests/test_audio_features.py:108-116 through the manager (rms_energy column)."""
    sr, amp = 22050, 0.7
    t = np.linspace(0, 1.0, sr, endpoint=False)
    r = extract_features(amp * np.sin(2 * np.pi * 440.0 * t), sr, ["rms_energy"], 1024, 512, output_format="dict_of_arrays")
    v = r["rms_energy"]
    assert v.ndim == 1 and v.dtype == np.float64 and np.all(v >= 0)
    assert_allclose(np.mean(v), amp / np.sqrt(2), atol=0.05)
    z = extract_features(np.zeros(sr), sr, ["rms_energy"], 1024, 512, output_format="dict_of_arrays")["rms_energy"]
    assert_allclose(z, 0.0, atol=1e-7)


def test_audio_features_module_mirror(engine):
    """tests/test_audio_features.py:95-116 against the mirror of sygnals.core.audio.features (same names and signatures)."""
    from sygnals_b200.core.audio.features import rms_energy, zero_crossing_rate
    sr, freq, amp = 22050, 440.0, 0.7
    t = np.linspace(0, 1.0, sr, endpoint=False)
    sine = (amp * np.sin(2 * np.pi * freq * t)).astype(np.float64)
    zcr = zero_crossing_rate(sine, frame_length=1024, hop_length=512)
    assert zcr.ndim == 1 and zcr.dtype == np.float64 and np.all(zcr >= 0) and len(zcr) == 1 + sr // 512
    assert abs(np.mean(zcr) - 2 * freq / sr) < 0.05
    assert_allclose(zero_crossing_rate(np.zeros(sr), frame_length=1024, hop_length=512), 0.0, atol=1e-7)
    assert np.mean(zero_crossing_rate(np.random.default_rng(0).standard_normal(sr), frame_length=1024, hop_length=512)) > 0.1
    rms = rms_energy(y=sine, frame_length=1024, hop_length=512)
    assert rms.ndim == 1 and rms.dtype == np.float64 and np.all(rms >= 0)
    assert_allclose(np.mean(rms), amp / np.sqrt(2), atol=0.05)
    assert_allclose(rms_energy(y=np.zeros(sr), frame_length=1024, hop_length=512), 0.0, atol=1e-7)
    with pytest.raises(ValueError):
        rms_energy()
    with pytest.raises(ValueError):
        rms_energy(y=np.zeros((2, 100)))
    with pytest.raises(NotImplementedError):
        rms_energy(S=np.ones((5, 4)))


# ------------------------------------------------------------------------------------------------ dsp
def test_compute_stft(engine):
    fs, freq = 1000.0, 50.0
    t = np.arange(0, 1.0, 1.0 / fs)
    x = np.sin(2 * np.pi * freq * t)
    D = dsp.compute_stft(x, n_fft=512, window="hann")
    assert D.shape[0] == 257 and D.dtype == np.complex128 and D.shape[1] == 1 + len(x) // 128
    bins = np.fft.rfftfreq(512, 1.0 / fs)
    assert np.argmax(np.mean(np.abs(D) ** 2, axis=1)) == np.argmin(np.abs(bins - freq))
    with pytest.raises(ValueError):
        dsp.compute_stft(np.zeros((2, 4096)))


# ------------------------------------------------------------------------------------------------ segmentation
@pytest.fixture
def ramp():
    sr = 1000
    return np.arange(int(sr * 5.3)).astype(np.float64) / sr, sr


def test_segment_fixed_length_no_overlap_no_pad(ramp):
    y, sr = ramp
    segs = segmentation.segment_fixed_length(y, sr, segment_length_sec=1.0, overlap_ratio=0.0, pad=False)
    assert isinstance(segs, list) and len(segs) == len(y) // 1000
    for i, s in enumerate(segs):
        assert isinstance(s, np.ndarray) and s.dtype == np.float64 and len(s) == 1000
        assert_array_equal(s, y[i * 1000:(i + 1) * 1000])


def test_segment_fixed_length_with_overlap_no_pad(ramp):
    y, sr = ramp
    segs = segmentation.segment_fixed_length(y, sr, segment_length_sec=1.0, overlap_ratio=0.5, pad=False)
    n, start = 0, 0
    while start + 1000 <= len(y):
        n += 1
        start += 500
    assert len(segs) == n
    for i, s in enumerate(segs):
        assert_array_equal(s, y[i * 500:i * 500 + 1000])


def test_segment_fixed_length_with_padding(ramp):
    y, sr = ramp
    segs = segmentation.segment_fixed_length(y, sr, segment_length_sec=1.0, overlap_ratio=0.25, pad=True)
    n, start = 0, 0
    while start < len(y):
        n += 1
        start += 750
    assert len(segs) == n
    last, s0 = segs[-1], (n - 1) * 750
    orig = len(y) - s0
    assert len(last) == 1000 and orig > 0
    assert_array_equal(last[:orig], y[s0:])
    assert_array_equal(last[orig:], np.zeros(1000 - orig))


def test_segment_fixed_length_min_length_and_short(ramp):
    y, sr = ramp
    segs = segmentation.segment_fixed_length(y, sr, segment_length_sec=1.0, overlap_ratio=0.0, pad=True, min_segment_length_sec=0.4)
    assert len(segs) == 5 and all(len(s) == 1000 for s in segs)
    short = y[:500]
    assert len(segmentation.segment_fixed_length(short, sr, 1.0, pad=False)) == 0
    p = segmentation.segment_fixed_length(short, sr, 1.0, pad=True)
    assert len(p) == 1 and len(p[0]) == 1000
    assert_array_equal(p[0][:500], short)
    assert_array_equal(p[0][500:], np.zeros(500))
    assert len(segmentation.segment_fixed_length(short, sr, 1.0, pad=True, min_segment_length_sec=0.6)) == 0


def test_segment_fixed_length_invalid_params(ramp):
    y, sr = ramp
    with pytest.raises(ValueError):
        segmentation.segment_fixed_length(y, sr, segment_length_sec=0)
    with pytest.raises(ValueError):
        segmentation.segment_fixed_length(y, sr, segment_length_sec=1.0, overlap_ratio=1.0)
    with pytest.raises(ValueError):
        segmentation.segment_fixed_length(y, sr, segment_length_sec=1.0, overlap_ratio=-0.1)
    with pytest.raises(ValueError):
        segmentation.segment_fixed_length(np.zeros((2, 10)), sr, 1.0)


@pytest.mark.gpu
def test_two_streams_share_one_engine_without_corrupting_the_workspace():
    """the device entry points are asynchronous on the caller's stream but share the context's workspace (mel energies,
    unit maxima).  Calls queued on two different torch streams must not overwrite each other's intermediates: the library orders
    them with an event (`ws_done`).  Results must equal the single-stream results bit for bit."""
    import torch
    from sygnals_b200 import _ffi
    from sygnals_b200.utils import synth
    eng = _ffi.engine(0)
    sr = 22050
    feats = ["mfcc", "spectral_contrast", "rms_energy"]
    p = _ffi.make_params(eng.lib, sr, feats, 2048, 512)
    ys = [torch.from_numpy(synth.clip_batch(300, 3 * sr, sr, seed=40 + i, edges=False)).cuda() for i in range(2)]
    units = eng.units_clips(300, 3 * sr)
    rows, T = eng.rows(p), eng.frame_count(3 * sr, 2048, 512, True)
    ref = []
    for y in ys:
        o = torch.empty((300, rows, T), dtype=torch.float32, device="cuda")
        eng.features_dev(y.data_ptr(), units, p, o.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ref.append(o)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [torch.empty_like(ref[0]) for _ in range(2)]
    torch.cuda.synchronize()
    for rep in range(3):                                             # interleave the two streams several times
        for i, st in enumerate(streams):
            eng.features_dev(ys[i].data_ptr(), units, p, outs[i].data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    for i in range(2):
        assert torch.equal(outs[i], ref[i])
