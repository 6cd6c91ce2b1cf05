"""BASELINE.json's full-size configurations on the B200 (``-m gpu``): exact counts, size-independent properties (idempotence,
batch independence, host path == device path, checksum-of-checksums) and oracle parity on sampled units.  Inputs are generated on
the device (sygnals_b200.utils.synth.torch_mixture_); sampled units are copied back for the float64 oracle."""
import numpy as np
import pytest

from test_parity_cabi import check_rows, oracle_rows

pytestmark = pytest.mark.gpu

CFG4_FEATURES = ["mfcc", "spectral_contrast", "spectral_centroid", "spectral_rolloff", "rms_energy", "crest_factor"]


def _torch():
    import torch
    assert torch.cuda.is_available()
    return torch


def test_cfg4_ten_hours_env_sound():
    torch = _torch()
    from sygnals_b200 import batch
    from sygnals_b200.utils import synth
    sr, total = 44100, 10 * 3600 * 44100
    y = torch.empty(total, dtype=torch.float32, device="cuda")
    synth.torch_mixture_(y, sr, seed=99, unit=sr)
    r = batch.segment_features(y, sr, 2.0, CFG4_FEATURES, overlap_ratio=0.5, pad=True)
    out = r["features"]
    torch.cuda.synchronize()
    assert tuple(out.shape) == (36000, 24, 173) and r["seg_len"] == 88200 and r["seg_hop"] == 44100
    assert int(r["valid"][-1]) == 44100 and int(r["starts"][-1]) == 35999 * 44100                       # last segment: half zeros
    assert bool(torch.isfinite(out).all())
    # idempotence: a second pass over the resident buffer reproduces every bit
    out2 = batch.segment_features(y, sr, 2.0, CFG4_FEATURES, overlap_ratio=0.5, pad=True)["features"]
    assert torch.equal(out, out2)
    # unit independence: any sub-range of the recording gives the same rows for its segments (per-unit ref=np.max)
    u0 = 12345
    sub = batch.segment_features(y[u0 * 44100:(u0 + 9) * 44100 + 88200], sr, 2.0, CFG4_FEATURES, overlap_ratio=0.5, pad=False)["features"]
    assert torch.equal(sub[:8], out[u0:u0 + 8])
    # host path (C ABI with host buffers) == device path on a slice, bit for bit
    yh = y[: 40 * 44100 + 44100].cpu().numpy()
    rh = batch.segment_features(yh, sr, 2.0, CFG4_FEATURES, overlap_ratio=0.5, pad=False)["features"]
    assert np.array_equal(rh, out[: rh.shape[0]].cpu().numpy())
    # oracle parity on sampled segments (first, last = zero padded tail, and a spread)
    rng = np.random.default_rng(4)
    for u in [0, 35999] + sorted(rng.integers(1, 35999, size=6).tolist()):
        s, v = int(r["starts"][u]), int(r["valid"][u])
        seg = np.zeros(88200, dtype=np.float32)
        seg[:v] = y[s:s + v].cpu().numpy()
        names, ref = oracle_rows(seg, sr, CFG4_FEATURES, 2048, 512)
        assert names == r["names"]
        S = np.abs(np.fft.rfft(np.lib.stride_tricks.sliding_window_view(np.pad(seg.astype(np.float64), 1024), 2048)[::512]
                               * np.hanning(2049)[:-1], axis=1)).T
        ok = S.min(axis=0) >= 1e-5 * S.max(axis=0)
        check_rows(names, out[u].cpu().numpy(), ref, bin_hz=sr / 2048, contrast_ok=ok)


def test_cfg3_speech_commands_100k_clips():
    torch = _torch()
    from sygnals_b200 import batch
    from sygnals_b200.utils import synth
    sr, n, L = 16000, 100000, 16000
    clips = torch.empty((n, L), dtype=torch.float32, device="cuda")
    synth.torch_mixture_(clips, sr, seed=7)
    fp = {"mfcc": {"n_mels": 40}}
    names, out = batch.extract_features_batch(clips, sr, ["mfcc"], 512, 160, feature_params=fp)
    torch.cuda.synchronize()
    assert tuple(out.shape) == (n, 13, 101) and names == [f"mfcc_{i}" for i in range(13)]
    assert bool(torch.isfinite(out).all())
    idx = [0, 1, 31, 4097, 50000, 99999]
    _, sub = batch.extract_features_batch(clips[idx].contiguous(), sr, ["mfcc"], 512, 160, feature_params=fp)
    assert torch.equal(sub, out[idx])                                   # clip i does not depend on its batch
    for i in idx[:4]:
        nm, ref = oracle_rows(clips[i].cpu().numpy(), sr, ["mfcc"], 512, 160, fp)
        check_rows(nm, out[i].cpu().numpy(), ref)


@pytest.mark.parametrize("n_fft", [256, 512, 1024, 2048, 4096, 8192])
def test_cfg2_stft_sweep_4096_clips(n_fft):
    torch = _torch()
    from oracle import sygnals_oracle as orc
    from sygnals_b200 import batch
    from sygnals_b200.utils import synth
    sr, n, L = 16000, 4096, 16000
    clips = torch.empty((n, L), dtype=torch.float32, device="cuda")
    synth.torch_mixture_(clips, sr, seed=n_fft)
    mag = batch.stft_batch(clips, n_fft=n_fft, output="magnitude")
    torch.cuda.synchronize()
    B, T = 1 + n_fft // 2, 1 + L // (n_fft // 4)
    assert tuple(mag.shape) == (n, B, T)
    pw = batch.stft_batch(clips, n_fft=n_fft, output="power")
    assert torch.allclose(mag * mag, pw, rtol=1e-5, atol=1e-12)
    for i in (0, 2048, 4095):
        ref = np.abs(orc.compute_stft(clips[i].cpu().numpy().astype(np.float64), n_fft=n_fft)) ** 2
        P = pw[i].cpu().numpy().astype(np.float64)
        tol = 1e-4 * ref + 1e-6 * ref.max(axis=0, keepdims=True)
        assert (np.abs(P - ref) <= tol).all()


def test_cfg5_machinery_64_channels():
    torch = _torch()
    import scipy.signal
    from sygnals_b200 import batch
    from sygnals_b200.utils import synth
    fs, ch, seconds = 25600, 64, 120
    y = torch.empty((ch * seconds, fs), dtype=torch.float32, device="cuda")       # unit = (channel, 1 s window)
    synth.torch_mixture_(y, fs, seed=55)
    psd, st = batch.psd_welch_batch(y, fs, nperseg=1024, noverlap=512)
    torch.cuda.synchronize()
    assert tuple(psd.shape) == (ch * seconds, 513) and tuple(st.shape) == (ch * seconds, 3)
    for u in (0, 777, ch * seconds - 1):
        x = y[u].cpu().numpy().astype(np.float64)
        _, ref = scipy.signal.welch(x, fs=fs, window="hann", nperseg=1024, noverlap=512, detrend="constant", scaling="density")
        P = psd[u].cpu().numpy().astype(np.float64)
        assert (np.abs(P - ref) <= 1e-4 * ref + 1e-6 * ref.max()).all()
        rms = np.sqrt(np.mean(x * x))
        np.testing.assert_allclose(st[u, 0].item(), rms, rtol=1e-5)
        np.testing.assert_allclose(st[u, 1].item(), np.abs(x).max() / rms, rtol=1e-5)


def test_segment_aggregation_on_device_and_mirror():
    torch = _torch()
    from oracle import sygnals_oracle as orc
    from sygnals_b200 import batch, dist as sdist
    from sygnals_b200.core.ml_utils import formatters as F
    from sygnals_b200.utils import synth
    sr = 44100
    y = torch.empty(600 * sr, dtype=torch.float32, device="cuda")
    synth.torch_mixture_(y, sr, seed=3, unit=sr)
    r = batch.segment_features(y, sr, 2.0, CFG4_FEATURES, overlap_ratio=0.5)
    feats, names = r["features"], r["names"]
    agg = {"mfcc_0": "median", "spectral_centroid": "std", "rms_energy": "max", "crest_factor": "min"}
    dev = F.aggregate_segments(feats, names, agg)
    torch.cuda.synchronize()
    assert tuple(dev.shape) == (feats.shape[0], 24) and dev.dtype == torch.float64
    h = feats.cpu().numpy().astype(np.float64)
    for s in (0, 17, feats.shape[0] - 1):
        ref = orc.format_feature_vectors_per_segment({n: h[s, j] for j, n in enumerate(names)}, [(0, h.shape[2])], agg)[0]
        np.testing.assert_allclose(dev[s].cpu().numpy(), ref, rtol=1e-12, atol=1e-12)
    # sharded entry point with on-device aggregation (single rank here: no process group) == aggregate of the plain result
    sh = sdist.segment_features_sharded(y, sr, 2.0, CFG4_FEATURES, overlap_ratio=0.5, aggregation=agg, rank=0, world=1)
    assert torch.equal(sh["features"], dev)
    # drop-in mirror of the reference function on a dict of 1-D arrays with (start, end) frame tables
    d = {n: h[5, j] for j, n in enumerate(names[:6])}
    segs = [(0, 50), (50, 173), (10, 11), (170, 180), (0, 173)]
    import warnings
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        got = F.format_feature_vectors_per_segment(d, segs, aggregation="median", output_format="numpy")
        df = F.format_feature_vectors_per_segment(d, segs, aggregation="mean")
    assert any("Invalid segment indices" in str(x.message) for x in w)
    ref = orc.format_feature_vectors_per_segment(d, segs, "median")
    assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.array_equal(got[~np.isnan(ref)], ref[~np.isnan(ref)])
    assert list(df.columns) == names[:6] and df.index.name == "segment_index" and np.isnan(df.iloc[3]).all()
