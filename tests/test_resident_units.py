"""Unit-resident feature kernel (frame_warp_kernel STAGE 5, syg_frame_warp.cuh): short units whose mel tile, power_to_db(ref=max,
top_db) and DCT stay inside the frame kernel (BASELINE cfg3 shape).  Reference: sygnals/core/features/manager.py:205-227 (per-call
``ref=np.max``), sygnals/core/features/cepstral.py:94-117.  Checked against the golden vectors of the unmodified reference, the
oracle, and -- bit for bit on everything except the DCT's summation order -- the two-kernel path."""
import numpy as np
import pytest

import cases
from backends import BACKENDS, get_engine
from sygnals_b200 import _ffi
from sygnals_b200.utils import synth
from test_parity_cabi import check_rows, oracle_rows


@pytest.fixture(params=BACKENDS)
def eng(request):
    e = get_engine(request.param)
    e.lib.dll.syg_debug_set_resident_min_groups(1)
    yield e
    e.lib.dll.syg_debug_set_resident_min_groups(-1)


def test_golden_cfg3_on_the_resident_kernel(eng):
    y, sr = cases.cfg3_input()
    g = cases.load("cfg3_speech_mfcc.npz")
    p = _ffi.make_params(eng.lib, sr, ["mfcc"], 512, 160, feature_params={"mfcc": {"n_mels": 40}})
    out = eng.features_host(y.ravel(), eng.units_clips(y.shape[0], y.shape[1]), p)
    assert eng.lib.dll.syg_debug_last_features_resident() == 1
    assert out.shape == g["rows"].shape
    check_rows([str(n) for n in g["names"]], out, g["rows"])


@pytest.mark.parametrize("n_fft,hop,L,n,feats", [(512, 160, 16000, 7, ["mfcc"]), (512, 160, 16000, 11, ["mfcc", "rms_energy", "spectral_centroid", "spectral_rolloff"]),
                                                 (256, 64, 3001, 5, ["mfcc", "crest_factor"]), (1024, 256, 9000, 6, ["rms_energy", "mfcc"]),
                                                 (512, 100, 1234, 13, ["mfcc"])])
def test_resident_equals_two_kernel_path_and_oracle(eng, n_fft, hop, L, n, feats):
    """Group sizes that do not divide the unit count, frame counts that do not fill the last round of warp tasks, silent and
    full-scale units inside a group (each unit keeps its own reference level)."""
    sr = 16000
    y = np.stack([synth.long_signal(L, sr, seed=100 * n_fft + c).astype(np.float32) * (10.0 ** (-c)) for c in range(n)])
    y[2] = 0.0
    fp = {"mfcc": {"n_mels": 40, "n_mfcc": 13}}
    p = _ffi.make_params(eng.lib, sr, feats, n_fft, hop, feature_params=fp)
    u = eng.units_clips(n, L)
    res = eng.features_host(y.ravel(), u, p)
    assert eng.lib.dll.syg_debug_last_features_resident() == 1
    eng.lib.dll.syg_debug_set_resident_min_groups(1 << 30)              # the two-kernel path
    two = eng.features_host(y.ravel(), u, p)
    assert eng.lib.dll.syg_debug_last_features_resident() == 0
    eng.lib.dll.syg_debug_set_resident_min_groups(1)
    assert res.shape == two.shape
    np.testing.assert_allclose(res, two, rtol=0, atol=2e-5)           # same energies, same dB; the DMMA tiles sum in the same order
    for c in range(n):
        names, ref = oracle_rows(y[c], sr, feats, n_fft, hop, feature_params=fp)
        check_rows(names, res[c], ref, bin_hz=sr / n_fft, nyq=sr / 2)


def test_requests_outside_the_resident_kernel_take_the_two_kernel_path(eng):
    sr = 16000
    y = synth.clip_batch(4, 4000, sr, seed=5, edges=False)
    u = eng.units_clips(4, 4000)
    p = _ffi.make_params(eng.lib, sr, ["mfcc", "spectral_contrast"], 512, 160, feature_params={"mfcc": {"n_mels": 40}})
    eng.features_host(y.ravel(), u, p)
    assert eng.lib.dll.syg_debug_last_features_resident() == 0          # contrast needs the unit-wide peak / valley maxima
    p = _ffi.make_params(eng.lib, sr, ["mfcc"], 2048, 512, feature_params={"mfcc": {"n_mels": 128}})
    eng.features_host(y.ravel(), u, p)
    assert eng.lib.dll.syg_debug_last_features_resident() == 0          # one frame per warp: the n_fft 2048 kernel keeps its workspace
