"""The sygnals-b200 plugin against the reference's own plugin machinery (build container: the UNMODIFIED reference loader /
registry from /root/reference) and, on the GPU box, the routed callables against the oracle."""
import logging
import os

import numpy as np
import pytest

from oracle import ref_loader
from sygnals_b200 import plugin as plg

needs_ref = pytest.mark.skipif(not ref_loader.reference_available(), reason="/root/reference not present")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_manifest_matches_class():
    import tomllib
    with open(os.path.join(ROOT, "sygnals_b200", "plugin.toml"), "rb") as f:
        man = tomllib.load(f)
    p = plg.SygnalsB200Plugin()
    assert man["name"] == p.name == "sygnals-b200"
    assert man["version"] == p.version
    mod, cls = man["entry_point"].split(":")
    assert mod == "sygnals_b200.plugin" and getattr(plg, cls) is plg.SygnalsB200Plugin
    with open(os.path.join(ROOT, "pyproject.toml"), "rb") as f:
        proj = tomllib.load(f)
    assert proj["project"]["entry-points"]["sygnals.plugins"]["sygnals-b200"] == man["entry_point"]


@needs_ref
def test_reference_loader_accepts_the_plugin_and_rebinding_round_trips():
    logging.disable(logging.CRITICAL)
    try:
        ref_loader.load_reference()
        from pathlib import Path
        from sygnals.plugins import loader as L
        from sygnals.plugins.api import PluginRegistry, SygnalsPluginBase
        from sygnals.config.models import SygnalsConfig
        from sygnals.version import __version__ as core_version
        import importlib
        importlib.reload(plg)                                   # pick up the real SygnalsPluginBase
        assert issubclass(plg.SygnalsB200Plugin, SygnalsPluginBase)
        mpath = Path(ROOT) / "sygnals_b200" / "plugin.toml"
        man = L._parse_manifest(mpath)                          # loader.py:63-83: required keys present
        assert man is not None
        assert L._check_compatibility(man["name"], man["version"], man["sygnals_api"], core_version)
        import sygnals.cli.features_cmd as fc
        import sygnals.cli.segment_cmd as sc
        import sygnals.core.dsp as dsp
        import sygnals.core.features.manager as mgr
        orig = (mgr.extract_features, fc.extract_features, dsp.compute_stft, sc.segment_fixed_length)
        reg = PluginRegistry()
        ld = L.PluginLoader(SygnalsConfig(), reg)
        ld.plugin_manifests[man["name"]] = (man, None)          # as discover_and_load() records an entry-point plugin
        ld.plugin_sources[man["name"]] = "entry_point"
        ld._load_and_register(man["name"])                      # loader.py:210-288: setup + 9 register hooks
        assert "sygnals-b200" in ld.loaded_plugins
        inst = ld.loaded_plugins["sygnals-b200"]
        assert "b200_mfcc" in reg.list_features() and "b200_stft" in reg.list_transforms()
        # the imported copies the CLI calls are rebound, not just the defining modules
        assert fc.extract_features is mgr.extract_features and fc.extract_features is not orig[1]
        assert fc.extract_features.__wrapped_reference__ is orig[0]
        assert dsp.compute_stft is not orig[2] and sc.segment_fixed_length is not orig[3]
        # unsupported features stay on the reference path (routed to the ORIGINAL function, bit-identical result)
        y = np.sin(np.arange(4000) * 0.05)
        a = fc.extract_features(y, 22050, ["zero_crossing_rate"], frame_length=502, hop_length=125, output_format="dict_of_arrays")
        b = orig[0](y, 22050, ["zero_crossing_rate"], frame_length=502, hop_length=125, output_format="dict_of_arrays")
        np.testing.assert_array_equal(a["zero_crossing_rate"], b["zero_crossing_rate"])
        a = dsp.compute_stft(y, n_fft=298)                      # 2 * 149: a prime factor above 13 -> reference
        np.testing.assert_array_equal(a, orig[2](y, n_fft=298))
        # array-level audio features: rebound, unsupported geometry routed to the ORIGINAL (bit-identical)
        import sygnals.core.audio.features as af
        assert getattr(af.rms_energy, "__wrapped_reference__", None) is not None
        r_a = af.rms_energy(y=y, frame_length=502, hop_length=125)
        r_b = af.rms_energy.__wrapped_reference__(y=y, frame_length=502, hop_length=125)
        np.testing.assert_array_equal(r_a, r_b)
        z_a = af.zero_crossing_rate(y, frame_length=502, hop_length=125)
        np.testing.assert_array_equal(z_a, af.zero_crossing_rate.__wrapped_reference__(y, frame_length=502, hop_length=125))
        # error contract unchanged (manager.py:141-143)
        with pytest.raises(ValueError, match="Unknown feature"):
            fc.extract_features(y, 22050, ["nope"])
        # segmentation is integer work: the engine's table must reproduce the reference list exactly (no GPU needed)
        segs_a = sc.segment_fixed_length(y, 1000, 1.0, overlap_ratio=0.5)
        segs_b = orig[3](y, 1000, 1.0, overlap_ratio=0.5)
        assert len(segs_a) == len(segs_b) and all(np.array_equal(p, q) for p, q in zip(segs_a, segs_b))
        inst.teardown()
        assert (mgr.extract_features, fc.extract_features, dsp.compute_stft, sc.segment_fixed_length) == orig
        assert getattr(af.rms_energy, "__wrapped_reference__", None) is None
    finally:
        logging.disable(logging.NOTSET)


def test_strict_mode_refuses_unsupported_instead_of_falling_back():
    p = plg.SygnalsB200Plugin()
    p._strict = True
    fn = p.make_extract_features(original=lambda *a, **k: pytest.fail("must not reach the reference"))
    with pytest.raises(NotImplementedError, match="no CUDA kernel"):
        fn(np.zeros(4096), 22050, ["jitter"])
    with pytest.raises(NotImplementedError, match="prime factors"):
        fn(np.zeros(4096), 22050, ["mfcc"], frame_length=1006)     # 2 * 503


@pytest.mark.gpu
def test_routed_extract_features_matches_oracle_on_gpu():
    from oracle import sygnals_oracle as orc
    from sygnals_b200.utils import synth
    p = plg.SygnalsB200Plugin()
    fn = p.make_extract_features(original=None)
    sr = 22050
    y = synth.mixture(3 * sr, sr, seed=7).astype(np.float64)
    feats = ["mfcc", "rms_energy", "spectral_centroid"]
    got = fn(y, sr, feats, output_format="dict_of_arrays")
    ref = orc.extract_features(y, sr, feats)
    assert list(got) == list(ref)
    np.testing.assert_array_equal(got["time"], ref["time"])
    for k in ref:
        if k.startswith("mfcc"):
            np.testing.assert_allclose(got[k], ref[k], atol=1e-3, rtol=0)         # abs 1e-3 on MFCC (north_star)
        elif k != "time":
            np.testing.assert_allclose(got[k], ref[k], rtol=1e-5, atol=1e-7)
    df = fn(y, sr, feats)                                                         # DataFrame contract (manager.py:430-442)
    assert df.index.name == "time" and list(df.columns) == [k for k in ref if k != "time"]


def test_mixed_request_is_split_and_paths_are_logged(caplog):
    """Kernelled features stay on the engine, the others go to the reference, columns come back in the requested order; every
    hand-over to the reference is a WARNING naming the path (once per reason).  Engine and reference are stand-ins here."""
    p = plg.SygnalsB200Plugin()
    calls = {}

    def fake_ref(y, sr, features, **kw):
        calls["ref"] = list(features)
        out = {"time": np.arange(3.0)}
        out.update({f: np.full(3, 7.0) for f in features})
        return out

    import sygnals_b200.core.features.manager as m
    real = m.extract_features

    def fake_engine(y, sr, features, **kw):
        calls["eng"] = list(features)
        assert kw["output_format"] == "dict_of_arrays"
        out = {"time": np.arange(3.0)}
        for f in features:
            for col in ([f"mfcc_{i}" for i in range(2)] if f == "mfcc" else [f]):
                out[col] = np.full(3, 1.0)
        return out

    m.extract_features = fake_engine
    try:
        fn = p.make_extract_features(original=fake_ref)
        with caplog.at_level(logging.WARNING, logger="sygnals_b200.plugin"):
            got = fn(np.zeros(4096), 22050, ["jitter", "mfcc", "rms_energy", "hnr"], feature_params={"mfcc": {"n_mfcc": 2}},
                     output_format="dict_of_arrays")
            fn(np.zeros(4096), 22050, ["jitter", "mfcc"], feature_params={"mfcc": {"n_mfcc": 2}}, output_format="dict_of_arrays")
        assert calls["eng"] == ["mfcc"] and list(got) == ["time", "jitter", "mfcc_0", "mfcc_1", "rms_energy", "hnr"]
        assert got["jitter"][0] == 7.0 and got["mfcc_1"][0] == 1.0
        warned = [r.message for r in caplog.records if "REFERENCE (CPU) implementation" in r.message]
        assert len(warned) == 2 and "no CUDA kernel for ['jitter', 'hnr']" in warned[0]      # two distinct reasons, each once
        df = fn(np.zeros(4096), 22050, ["jitter", "mfcc"], feature_params={"mfcc": {"n_mfcc": 2}})
        assert df.index.name == "time" and list(df.columns) == ["jitter", "mfcc_0", "mfcc_1"]
        # a NotImplementedError raised by the engine after the routing checks also lands on the reference (non-strict)
        def refusing(y, sr, features, **kw):
            raise NotImplementedError("n_mels=300: supported range is [1, 256]")
        m.extract_features = refusing
        calls.clear()
        fn(np.zeros(4096), 22050, ["mfcc"], output_format="dict_of_arrays")
        assert calls["ref"] == ["mfcc"]
        p._strict = True
        with pytest.raises(NotImplementedError):
            p.make_extract_features(original=fake_ref)(np.zeros(4096), 22050, ["mfcc"], output_format="dict_of_arrays")
    finally:
        m.extract_features = real


def test_psd_wrappers_accept_positional_arguments():
    """compute_psd_welch(x, fs, window, nperseg, noverlap, nfft, ...) positionally, as the reference signature allows (dsp.py:495)."""
    p = plg.SygnalsB200Plugin()
    seen = {}

    def ref(x, fs=1.0, window="hann", **kw):
        seen.update(kw)
        return "ref"

    fn = p._make_psd("compute_psd_welch", ref)
    assert fn(np.zeros(3000), 1000.0, "hann", 1006, 500) == "ref"        # nperseg 1006 = 2 * 503: no kernel -> reference, same arguments
    assert seen == {"nperseg": 1006, "noverlap": 500}
    with pytest.raises(TypeError):
        fn(np.zeros(3000), 1000.0, "hann", 1006, nperseg=512)
