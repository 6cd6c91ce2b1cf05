"""The C-ABI library loads without a GPU and exports every symbol include/sygb200.h declares (no compute calls here); the
host-only entry points (frame count, segment table, plan tables) are checked against the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import sygnals_oracle as orc
from sygnals_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sygb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(syg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    syms = declared_symbols()
    assert len(syms) >= 25, syms
    lib = _ffi.library()                                   # the product library (nvcc-built, in-tree)
    dll = ctypes.CDLL(lib.path)
    for s in syms:
        assert hasattr(dll, s), f"{s} declared in include/sygb200.h but not exported by {lib.path}"
    assert set(syms) == set(lib.symbols), set(syms) ^ set(lib.symbols)      # the ctypes binding covers exactly the header
    assert "sm_100a" in lib.version()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(_ffi.EngineError, match="no CPU fallback|no CUDA device"):
        _ffi.Engine(0)


@pytest.mark.parametrize("n,fl,hop,center", [(220500, 2048, 512, True), (16000, 512, 160, True), (100, 1024, 512, True),
                                             (5000, 1024, 300, False), (1023, 1024, 256, False), (0, 512, 128, True)])
def test_frame_count_matches_reference_rule(n, fl, hop, center):
    lib = _ffi.library()
    want = (1 + n // hop) if center else (1 + (n - fl) // hop if n >= fl else 0)     # manager.py:149-157 (even frame_length)
    assert lib.frame_count(n, fl, hop, center) == want


@pytest.mark.parametrize("total,sr,sec,ovl,pad,mn", [(5300, 1000, 1.0, 0.25, True, None), (5300, 1000, 1.0, 0.5, False, None),
                                                     (5300, 1000, 0.7, 0.3, True, 0.2), (88200 * 3 + 17, 44100, 2.0, 0.5, True, None),
                                                     (10, 1000, 1.0, 0.0, True, None), (0, 1000, 1.0, 0.0, True, None)])
def test_segment_table_is_bit_exact(total, sr, sec, ovl, pad, mn):
    lib = _ffi.library()
    y = np.arange(total, dtype=np.float64)
    segs = orc.segment_fixed_length(y, sr, sec, overlap_ratio=ovl, pad=pad, min_segment_length_sec=mn)
    seg_len, seg_hop, starts, valid = lib.segment_table(total, sr, sec, ovl, pad, mn)
    assert len(segs) == len(starts)
    for s, st, v in zip(segs, starts, valid):
        assert len(s) == seg_len
        np.testing.assert_array_equal(s[:v], y[st:st + v])
        assert not s[v:].any()


def test_plan_tables_match_oracle_shim():
    from oracle import librosa_shim as L
    lib = _ffi.library()
    np.testing.assert_allclose(lib.debug_mel_basis(44100, 2048, 128), L.filters.mel(sr=44100, n_fft=2048, n_mels=128), atol=2e-7)
    np.testing.assert_allclose(lib.debug_mel_basis(16000, 512, 40), L.filters.mel(sr=16000, n_fft=512, n_mels=40), atol=2e-7)
    import scipy.fftpack
    np.testing.assert_allclose(lib.debug_dct(13, 40), scipy.fftpack.dct(np.eye(40), type=2, norm="ortho", axis=0)[:13], atol=1e-7)
    import scipy.signal
    np.testing.assert_allclose(lib.debug_window(0, 2048, 2048), scipy.signal.get_window("hann", 2048, fftbins=True), atol=1e-7)


def test_numa_helper_is_a_noop_without_a_gpu():
    from sygnals_b200.utils import numa
    assert numa._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert numa._parse_cpulist("") == set()
    import os
    before = os.sched_getaffinity(0)
    r = numa.bind_to_device_node(0)
    assert r is None or set(os.sched_getaffinity(0)) <= set(before)
    os.sched_setaffinity(0, before)
