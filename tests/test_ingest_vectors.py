"""Ingest (load_audio on the device) and segment vectors (8f-1 fused behind the host entry points).

Reference: sygnals/core/audio/io.py:38-102 (load_audio -> librosa.load -> soundfile float32 read + to_mono),
sygnals/core/ml_utils/formatters.py:51-163 (format_feature_vectors_per_segment), sygnals/cli/save_cmd.py:140-190.
The CUDA path runs through the C ABI; the oracle (oracle/sygnals_oracle.py) is the checker.
"""
import os
import struct
import wave

import numpy as np
import pytest

from backends import BACKENDS, get_engine
from oracle import sygnals_oracle as orc
from sygnals_b200 import _ffi
from sygnals_b200.core.audio import io as sio
from sygnals_b200.utils import synth


@pytest.fixture(params=BACKENDS)
def eng(request):
    return get_engine(request.param)


def _dev(eng, *arrays):
    if getattr(eng, "test_backend", None) == "emu":
        hold = [np.ascontiguousarray(a) for a in arrays]
        return hold, [h.ctypes.data for h in hold]
    import torch
    hold = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrays]
    return hold, [h.data_ptr() for h in hold]


def _host(h):
    if isinstance(h, np.ndarray):
        return h
    import torch
    torch.cuda.synchronize()
    return h.cpu().numpy()


def _payload(fmt, channels, frames, seed):
    """Random interleaved payload bytes incl. the extreme codes of the format."""
    rng = np.random.default_rng(seed)
    n = frames * channels
    if fmt == _ffi.PCM_U8:
        x = rng.integers(0, 256, n).astype(np.uint8)
        x[:3] = np.array([0, 255, 128], dtype=np.uint8)[: x[:3].size]
        return x
    if fmt == _ffi.PCM_S16:
        x = rng.integers(-32768, 32768, n).astype("<i2")
        x[:4] = np.array([-32768, 32767, 0, -1])[: x[:4].size]
        return x.view(np.uint8)
    if fmt == _ffi.PCM_S24:
        v = rng.integers(-(1 << 23), 1 << 23, n).astype(np.int64)
        v[:4] = np.array([-(1 << 23), (1 << 23) - 1, 0, -1])[: v[:4].size]
        b = np.empty((n, 3), dtype=np.uint8)
        b[:, 0], b[:, 1], b[:, 2] = v & 255, (v >> 8) & 255, (v >> 16) & 255
        return b.reshape(-1)
    if fmt == _ffi.PCM_S32:
        x = rng.integers(-(1 << 31), 1 << 31, n).astype("<i4")
        x[:4] = np.array([-(1 << 31), (1 << 31) - 1, 0, -1])[: x[:4].size]
        return x.view(np.uint8)
    x = (rng.standard_normal(n) * 10.0 ** rng.integers(-4, 1, n)).astype("<f4")
    return x.view(np.uint8)


@pytest.mark.parametrize("fmt", [_ffi.PCM_U8, _ffi.PCM_S16, _ffi.PCM_S24, _ffi.PCM_S32, _ffi.PCM_F32])
@pytest.mark.parametrize("channels", [1, 2, 3, 6, 8, 11, 17])
def test_ingest_pcm_bit_exact(eng, fmt, channels):
    """syg_ingest_pcm == soundfile float32 read + np.mean over channels, bit for bit (ragged frame counts, empty payload)."""
    for frames in (0, 1, 7, 1000, 4099):
        raw = _payload(fmt, channels, frames, seed=fmt * 100 + channels)
        ref = orc.load_audio_payload(raw, fmt, channels, mono=True).astype(np.float32)
        out = np.full(max(frames, 1), 7.0, dtype=np.float32)
        hold, (praw, pout) = _dev(eng, raw if raw.size else np.zeros(16, np.uint8), out)
        eng.ingest_pcm_dev(praw, fmt, channels, frames, pout)
        got = _host(hold[1])[:frames]
        assert got.tobytes() == ref.tobytes() or np.array_equal(got, ref), (fmt, channels, frames)


def _write_wav(path, raw, sr, fmt, channels, extensible=False):
    bits = {_ffi.PCM_U8: 8, _ffi.PCM_S16: 16, _ffi.PCM_S24: 24, _ffi.PCM_S32: 32, _ffi.PCM_F32: 32}[fmt]
    tag = 3 if fmt == _ffi.PCM_F32 else 1
    ba = channels * bits // 8
    if extensible:
        guid = struct.pack("<H", tag) + bytes.fromhex("000000001000800000aa00389b71")
        fmt_body = struct.pack("<HHIIHHHHI", 0xFFFE, channels, sr, sr * ba, ba, bits, 22, bits, 0) + guid
    else:
        fmt_body = struct.pack("<HHIIHH", tag, channels, sr, sr * ba, ba, bits)
    data = bytes(raw)
    junk = b"LIST" + struct.pack("<I", 5) + b"abcde" + b"\x00"                  # an odd-sized chunk in front: word alignment
    body = b"WAVE" + junk + b"fmt " + struct.pack("<I", len(fmt_body)) + fmt_body + b"data" + struct.pack("<I", len(data)) + data
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", len(body)) + body)


@pytest.mark.parametrize("fmt,channels,ext", [(_ffi.PCM_S16, 1, False), (_ffi.PCM_S16, 2, True), (_ffi.PCM_S24, 2, False),
                                              (_ffi.PCM_S32, 1, True), (_ffi.PCM_F32, 2, False), (_ffi.PCM_U8, 1, False)])
def test_wav_header_parse(tmp_path, fmt, channels, ext):
    """Host-side RIFF parse: format, channel count, rate, payload offset / length, offset + duration arithmetic."""
    raw = _payload(fmt, channels, 3001, seed=5)
    p = tmp_path / "a.wav"
    _write_wav(p, raw, 22050, fmt, channels, extensible=ext)
    info = sio.wav_info(p)
    assert (info.sample_rate, info.channels, info.fmt, info.n_frames) == (22050, channels, fmt, 3001)
    mm, frames = sio.wav_payload(info)
    assert frames == 3001 and bytes(mm) == bytes(raw)
    mm2, frames2 = sio.wav_payload(info, offset=0.05, duration=0.02)             # librosa: int(offset * sr), int(duration * sr)
    fb = info.frame_bytes
    assert frames2 == int(0.02 * 22050) and bytes(mm2) == bytes(raw[int(0.05 * 22050) * fb:][: frames2 * fb])
    if fmt == _ffi.PCM_S16 and not ext:                                          # the stdlib reader agrees on a plain PCM16 file
        with wave.open(str(p), "rb") as w:
            assert (w.getnchannels(), w.getframerate(), w.getnframes()) == (channels, 22050, 3001)
    with pytest.raises(ValueError):
        bad = tmp_path / "b.wav"
        bad.write_bytes(b"RIFX" + bytes(40))
        sio.wav_info(bad)


@pytest.mark.gpu
def test_load_audio_wav_gpu(tmp_path):
    """load_audio mirror on the B200: float64, mono mean, (channels, n) for mono=False, FileNotFoundError like the reference."""
    raw = _payload(_ffi.PCM_S24, 2, 5000, seed=9)
    p = tmp_path / "s.wav"
    _write_wav(p, raw, 44100, _ffi.PCM_S24, 2)
    y, sr = sio.load_audio(p)
    assert sr == 44100 and y.dtype == np.float64 and y.shape == (5000,)
    assert np.array_equal(y, orc.load_audio_payload(raw, _ffi.PCM_S24, 2, mono=True))
    y2, _ = sio.load_audio(p, mono=False)
    assert y2.shape == (2, 5000) and np.array_equal(y2, orc.load_audio_payload(raw, _ffi.PCM_S24, 2, mono=False))
    with pytest.raises(FileNotFoundError):
        sio.load_audio(tmp_path / "missing.wav")
    with pytest.raises(NotImplementedError):
        sio.load_audio(p, sr=22050)


FEATS = ["mfcc", "spectral_centroid", "rms_energy", "crest_factor", "spectral_contrast", "spectral_rolloff"]


@pytest.mark.parametrize("fmt,channels", [(None, 1), (_ffi.PCM_S16, 1), (_ffi.PCM_S16, 2), (_ffi.PCM_S24, 3)])
def test_segment_vectors_host(eng, fmt, channels):
    """syg_segment_vectors_host_*: PCM payload in -> [segments, rows] float64 out.  Equals aggregation of the engine's own frame
    features bit for bit (same kernels, only the D2H differs) and the oracle's format_feature_vectors_per_segment within the
    frame-level tolerances."""
    sr, seg_sec = 22050, 0.5
    n = int(2.3 * sr)
    rng = np.random.default_rng(3)
    mono = synth.mixture(n, sr, seed=77)                                         # no silent / DC blocks: their contrast valleys are FP32 noise
    if fmt is None:
        raw, y32 = mono, mono
    else:
        chans = np.stack([mono * (0.5 + 0.5 * c) + 0.01 * rng.standard_normal(n).astype(np.float32) for c in range(channels)], axis=1)
        if fmt == _ffi.PCM_S16:
            raw = np.clip(np.round(chans * 20000), -32768, 32767).astype("<i2").view(np.uint8).reshape(-1)
        else:
            v = np.clip(np.round(chans * 5e6), -(1 << 23), (1 << 23) - 1).astype(np.int64).reshape(-1)
            b = np.empty((v.size, 3), dtype=np.uint8)
            b[:, 0], b[:, 1], b[:, 2] = v & 255, (v >> 8) & 255, (v >> 16) & 255
            raw = b.reshape(-1)
        y32 = orc.load_audio_payload(raw, fmt, channels).astype(np.float32)
    seg_len, seg_hop, starts, valid = eng.lib.segment_table(n, sr, seg_sec, 0.5, True, None)
    units = eng.units_clips(len(starts), seg_len, total_len=n, stride=seg_hop)
    fp = {"mfcc": {"n_mels": 64}}
    p = _ffi.make_params(eng.lib, sr, FEATS, 1024, 256, feature_params=fp)
    rows = eng.rows(p)
    aggs = ["mean", "std", "median", "min", "max"]
    ids = [_ffi.AGG_IDS[aggs[i % 5]] for i in range(rows)]
    raw = np.ascontiguousarray(raw)
    got = eng.segment_vectors_host(raw.ctypes.data, units, p, ids, fmt=fmt, channels=channels)
    assert got.shape == (len(starts), rows) and got.dtype == np.float64
    # (1) same as frame features + aggregate through the separate entry points
    frames = eng.features_host(y32, units, p)
    T = frames.shape[2]
    out = np.zeros((len(starts), rows), dtype=np.float64)
    hold, (pf, po) = _dev(eng, frames, out)
    eng.aggregate_dev(pf, len(starts), rows, T, ids, po, fixed_len=T)
    assert np.array_equal(_host(hold[1]), got, equal_nan=True)
    # (2) the oracle: reference call graph per segment, then format_feature_vectors_per_segment
    from sygnals_b200.batch import feature_row_names
    names = feature_row_names(FEATS, fp)
    segs = orc.segment_fixed_length(y32.astype(np.float64), sr, seg_sec, overlap_ratio=0.5, pad=True)
    assert len(segs) == len(starts)
    for i in (0, len(segs) // 2, len(segs) - 1):
        ref = orc.extract_features(segs[i], sr, FEATS, frame_length=1024, hop_length=256, feature_params=fp)
        d = {k: ref[k] for k in names}
        r = orc.format_feature_vectors_per_segment(d, [(0, T)], {k: aggs[j % 5] for j, k in enumerate(names)})[0]
        for j, k in enumerate(names):
            tol = 2e-3 if k.startswith("mfcc") else (3e-2 if k.startswith("contrast") else (sr / 1024 + 1e-6 if k == "spectral_rolloff" else 1e-4 * max(1.0, abs(r[j]))))
            assert abs(got[i, j] - r[j]) <= tol, (k, got[i, j], r[j])


def test_segment_vectors_device_equals_host(eng):
    """syg_segment_vectors_f32 (device buffers, chunked behind a small workspace limit) == the host form, bit for bit."""
    sr = 16000
    y = synth.long_signal(3 * sr + 123, sr, seed=8, block_sec=0.2)
    feats = ["mfcc", "rms_energy", "spectral_contrast"]
    p = _ffi.make_params(eng.lib, sr, feats, 512, 160, feature_params={"mfcc": {"n_mels": 40}})
    units = eng.units_clips(11, 8000, total_len=y.size, stride=4000)
    rows = eng.rows(p)
    ids = [_ffi.AGG_IDS["mean"]] * rows
    ref = eng.segment_vectors_host(y.ctypes.data, units, p, ids)
    out = np.zeros((11, rows), dtype=np.float64)
    eng.set_workspace_limit(1 << 20)                                             # several chunks
    try:
        hold, (py, po) = _dev(eng, y, out)
        eng.segment_vectors_dev(py, units, p, ids, po)
        got = _host(hold[1])
    finally:
        eng.set_workspace_limit(1 << 30)
    assert np.array_equal(got, ref, equal_nan=True)
    with pytest.raises(ValueError):
        eng.segment_vectors_host(y.ctypes.data, units, p, [9] * rows)            # unknown aggregation id
