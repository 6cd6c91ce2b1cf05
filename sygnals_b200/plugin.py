"""``sygnals-b200``: the Sygnals plugin that puts the B200 engine behind the reference's own names.

Follows the worked example ``plugins_contrib/sygnals-parquet`` of the reference (``plugin.toml:5-19``,
``sygnals_parquet/plugin.py:21-81``) against ``SygnalsPluginBase`` (``sygnals/plugins/api.py:37-138``); the loader calls
``setup(config.model_dump())`` then the nine ``register_*`` hooks (``sygnals/plugins/loader.py:255-274``) and
``teardown()`` at exit (``sygnals/cli/main.py:47-56``).

The registry alone is not enough: core never consults ``registry._features`` / ``_transforms`` (only readers/writers are
looked up, ``sygnals/core/data_handler.py:126,223``) and the CLI binds the hot-path functions *by imported name*
(``sygnals/cli/features_cmd.py:16``, ``sygnals/cli/segment_cmd.py:16``).  So ``setup()`` also REBINDS those module
attributes to the engine's mirrors (``sygnals_b200.core``) and ``teardown()`` restores them.

Routing rule (no CPU fallback inside the engine): a call is served by the engine when every requested feature has a CUDA
kernel and the frame geometry is supported; anything else (pitch/HNR/jitter/shimmer, non power-of-two
``frame_length``, exotic windows) is handed to the ORIGINAL reference function untouched -- those features "stay on the
reference path" exactly as the scope contract says.  Configure with ``[plugins.sygnals-b200]`` in the Sygnals
config: ``device`` (int, default LOCAL_RANK or 0), ``rebind`` (bool, default true), ``strict`` (bool, default false: raise
instead of routing unsupported calls to the reference).
"""
from __future__ import annotations

import importlib
import logging
from typing import Any, Callable, Dict, List, Optional, Tuple

from . import __version__, _ffi

logger = logging.getLogger(__name__)

try:                                        # the reference is only importable where Sygnals is installed
    from sygnals.plugins.api import SygnalsPluginBase  # type: ignore
except Exception:                           # pragma: no cover - exercised on boxes without Sygnals
    class SygnalsPluginBase:                # minimal stand-in with the same surface (api.py:37-138)
        def setup(self, config: Dict[str, Any]):
            pass

        def teardown(self):
            pass

PLUGIN_NAME = "sygnals-b200"

# (module that DEFINES the function, attribute, modules that hold an imported copy)
_REBIND_TARGETS: List[Tuple[str, str, Tuple[str, ...]]] = [
    ("sygnals.core.features.manager", "extract_features", ("sygnals.core.features", "sygnals.cli.features_cmd")),
    ("sygnals.core.dsp", "compute_stft", ("sygnals.core",)),
    ("sygnals.core.dsp", "compute_psd_welch", ("sygnals.core",)),
    ("sygnals.core.dsp", "compute_psd_periodogram", ("sygnals.core",)),
    ("sygnals.core.segmentation", "segment_fixed_length", ("sygnals.cli.segment_cmd",)),
    ("sygnals.core.audio.features", "rms_energy", ()),
    ("sygnals.core.audio.features", "zero_crossing_rate", ()),
]


def _pow2_in_range(n: int, lo: int = 32, hi: int = 8192) -> bool:
    return isinstance(n, int) and lo <= n <= hi and (n & (n - 1)) == 0


class SygnalsB200Plugin(SygnalsPluginBase):
    """Registers the engine's functions and rebinds the reference's hot-path names to them."""

    def __init__(self):
        self._saved: List[Tuple[Any, str, Any]] = []
        self._strict = False
        self._device: Optional[int] = None

    @property
    def name(self) -> str:
        return PLUGIN_NAME

    @property
    def version(self) -> str:
        return __version__

    # ------------------------------------------------------------------ lifecycle
    def setup(self, config: Dict[str, Any]):
        opts = ((config or {}).get("plugins") or {}).get(PLUGIN_NAME) or {}
        self._strict = bool(opts.get("strict", False))
        self._device = opts.get("device")
        lib = _ffi.library()                  # raises ImportError if libsygb200.so is missing: no silent CPU path
        logger.info(f"Initializing plugin '{self.name}' v{self.version}: {lib.version()} ({lib.path})")
        if opts.get("rebind", True):
            self.rebind()

    def teardown(self):
        self.restore()
        _ffi.shutdown()
        logger.info(f"Tearing down plugin '{self.name}'")

    # ------------------------------------------------------------------ routed callables
    def _engine_ok(self, what: str, reason: Optional[str], original: Optional[Callable]):
        """None if the engine serves the call; otherwise the original to delegate to (or raise when strict)."""
        if reason is None:
            return None
        if self._strict or original is None:
            raise NotImplementedError(f"{what}: {reason} (sygnals-b200 strict mode: not routed to the reference)")
        logger.debug(f"{what}: {reason}; served by the reference implementation")
        return original

    def make_extract_features(self, original: Optional[Callable]) -> Callable:
        from .core.features import manager as m

        def extract_features(y, sr, features, frame_length=2048, hop_length=512, center=True, window="hann",
                             feature_params=None, output_format="dataframe"):
            reason = None
            names = sorted(m._ALL_KNOWN_FEATURES) if features == ["all"] else list(features)
            missing = [f for f in names if f in m._ALL_KNOWN_FEATURES and f not in m.ENGINE_FEATURES]
            if missing:
                reason = f"no CUDA kernel for {missing}"
            elif not _pow2_in_range(frame_length):
                reason = f"frame_length={frame_length} is not a power of two in [32, 8192]"
            elif not isinstance(window, str) or window.lower() not in _ffi.WINDOW_IDS:
                reason = f"window={window!r} is not built into the engine"
            ref = self._engine_ok("extract_features", reason, original)
            fn = ref or m.extract_features
            return fn(y, sr, features, frame_length=frame_length, hop_length=hop_length, center=center, window=window,
                      feature_params=feature_params, output_format=output_format)

        extract_features.__doc__ = m.extract_features.__doc__
        extract_features.__wrapped_reference__ = original
        return extract_features

    def make_compute_stft(self, original: Optional[Callable]) -> Callable:
        from .core import dsp as d

        def compute_stft(y, n_fft=2048, hop_length=None, win_length=None, window="hann", center=True, pad_mode="constant"):
            reason = None
            if not _pow2_in_range(n_fft):
                reason = f"n_fft={n_fft} is not a power of two in [32, 8192]"
            elif not isinstance(window, str) or window.lower() not in _ffi.WINDOW_IDS:
                reason = f"window={window!r} is not built into the engine"
            elif pad_mode not in _ffi.PAD_IDS:
                reason = f"pad_mode={pad_mode!r} is not built into the engine"
            fn = self._engine_ok("compute_stft", reason, original) or d.compute_stft
            return fn(y, n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window, center=center,
                      pad_mode=pad_mode)

        compute_stft.__wrapped_reference__ = original
        return compute_stft

    def _make_psd(self, which: str, original: Optional[Callable]) -> Callable:
        from .core import dsp as d
        mirror = getattr(d, which)

        def psd(x, fs=1.0, window="hann", **kw):
            n = len(x) if hasattr(x, "__len__") else 0
            nfft = kw.get("nfft") or kw.get("nperseg") or (256 if which == "compute_psd_welch" else n)
            reason = None
            if not _pow2_in_range(int(nfft)):
                reason = f"nfft={nfft} is not a power of two in [32, 8192]"
            elif not isinstance(window, str) or window.lower() not in _ffi.WINDOW_IDS:
                reason = f"window={window!r} is not built into the engine"
            elif kw.get("detrend", "constant") not in ("constant", True, False):
                reason = f"detrend={kw.get('detrend')!r} is not built into the engine"
            fn = self._engine_ok(which, reason, original) or mirror
            return fn(x, fs=fs, window=window, **kw)

        psd.__name__ = which
        psd.__wrapped_reference__ = original
        return psd

    def make_segment_fixed_length(self, original: Optional[Callable]) -> Callable:
        from .core import segmentation as s

        def segment_fixed_length(y, sr, segment_length_sec, overlap_ratio=0.0, pad=True, min_segment_length_sec=None):
            return s.segment_fixed_length(y, sr, segment_length_sec, overlap_ratio=overlap_ratio, pad=pad,
                                          min_segment_length_sec=min_segment_length_sec)

        segment_fixed_length.__wrapped_reference__ = original
        return segment_fixed_length

    def _make_audio_feature(self, which: str, original: Optional[Callable]) -> Callable:
        from .core.audio import features as af
        mirror = getattr(af, which)

        def feature(y=None, *args, **kw):
            fl = args[0] if args else kw.get("frame_length", 2048)
            reason = None
            if not _pow2_in_range(fl):
                reason = f"frame_length={fl} is not a power of two in [32, 8192]"
            elif kw.get("S") is not None:
                reason = "spectrogram input has no CUDA kernel"
            elif kw.get("pad_mode", "constant") != "constant":
                reason = f"pad_mode={kw.get('pad_mode')!r} is not built into the engine"
            elif set(kw) - {"S", "frame_length", "hop_length", "center", "pad_mode"}:
                reason = "extra librosa keyword arguments"
            fn = self._engine_ok(which, reason, original) or mirror
            return fn(y, *args, **kw)

        feature.__name__ = which
        feature.__wrapped_reference__ = original
        return feature

    def _replacement(self, attr: str, original: Optional[Callable]) -> Callable:
        if attr == "extract_features":
            return self.make_extract_features(original)
        if attr == "compute_stft":
            return self.make_compute_stft(original)
        if attr in ("compute_psd_welch", "compute_psd_periodogram"):
            return self._make_psd(attr, original)
        if attr == "segment_fixed_length":
            return self.make_segment_fixed_length(original)
        if attr in ("rms_energy", "zero_crossing_rate"):
            return self._make_audio_feature(attr, original)
        raise KeyError(attr)

    # ------------------------------------------------------------------ rebinding
    def rebind(self) -> List[str]:
        """Point the reference's names at the engine.  Returns the dotted names that were rebound."""
        done: List[str] = []
        for mod_name, attr, copies in _REBIND_TARGETS:
            try:
                mod = importlib.import_module(mod_name)
            except Exception as e:            # Sygnals not installed / partially importable
                logger.debug(f"rebind: cannot import {mod_name}: {e}")
                continue
            original = getattr(mod, attr, None)
            if original is None or getattr(original, "__wrapped_reference__", None) is not None:
                continue
            new = self._replacement(attr, original)
            for target_name in (mod_name,) + tuple(copies):
                try:
                    target = importlib.import_module(target_name)
                except Exception:
                    continue
                if getattr(target, attr, None) is original:
                    self._saved.append((target, attr, original))
                    setattr(target, attr, new)
                    done.append(f"{target_name}.{attr}")
        if done:
            logger.info(f"'{self.name}' rebound: {', '.join(done)}")
        return done

    def restore(self) -> None:
        for target, attr, original in reversed(self._saved):
            setattr(target, attr, original)
        self._saved.clear()

    # ------------------------------------------------------------------ registry hooks (discoverability)
    def register_feature_extractors(self, registry):
        from . import batch

        def make(feature: str):
            def fn(y, sr, frame_length=2048, hop_length=512, center=True, window="hann", **params):
                """Frame-level feature of a 1-D signal on the B200 engine -> float64 [rows, T]."""
                import numpy as np
                y = np.asarray(y, dtype=np.float32).reshape(1, -1)
                _, out = batch.extract_features_batch(y, sr, [feature], frame_length, hop_length, center, window,
                                                      {feature: params} if params else None, device=self._device)
                out = out[0].astype(np.float64)
                return out[0] if out.shape[0] == 1 else out
            fn.__name__ = f"b200_{feature}"
            return fn

        for feature in sorted(_ffi.FEATURE_IDS):
            registry.add_feature(f"b200_{feature}", make(feature))

    def register_transforms(self, registry):
        from .core import dsp as d
        registry.add_transform("b200_stft", d.compute_stft)
        registry.add_transform("b200_psd_welch", d.compute_psd_welch)
        registry.add_transform("b200_psd_periodogram", d.compute_psd_periodogram)
