"""``sygnals-b200``: the Sygnals plugin that puts the B200 engine behind the reference's own names.

Follows the worked example ``plugins_contrib/sygnals-parquet`` of the reference (``plugin.toml:5-19``,
``sygnals_parquet/plugin.py:21-81``) against ``SygnalsPluginBase`` (``sygnals/plugins/api.py:37-138``); the loader calls
``setup(config.model_dump())`` then the nine ``register_*`` hooks (``sygnals/plugins/loader.py:255-274``) and
``teardown()`` at exit (``sygnals/cli/main.py:47-56``).

The registry alone is not enough: core never consults ``registry._features`` / ``_transforms`` (only readers/writers are
looked up, ``sygnals/core/data_handler.py:126,223``) and the CLI binds the hot-path functions *by imported name*
(``sygnals/cli/features_cmd.py:16``, ``sygnals/cli/segment_cmd.py:16``).  So ``setup()`` also REBINDS those module
attributes to the engine's mirrors (``sygnals_b200.core``) and ``teardown()`` restores them.

Routing rule (no CPU fallback inside the engine): a call is served by the engine when the frame geometry is supported and the
requested features have CUDA kernels.  A MIXED request (``['mfcc', 'jitter']``) is split: the kernelled features run on the engine
in one fused launch, only the others (pitch/HNR/jitter/shimmer) go to the ORIGINAL reference function, and the columns are merged
back in the requested order.  Calls the engine cannot serve at all (non power-of-two ``frame_length``, exotic windows, librosa
keyword arguments) are handed to the reference function untouched -- they "stay on the reference path" as the scope contract says
 -- and every such hand-over is logged at WARNING level, once per distinct reason, naming the path taken, so a user can
tell which implementation ran.  Configure with ``[plugins.sygnals-b200]`` in the Sygnals config: ``device`` (int, default LOCAL_RANK
or 0), ``rebind`` (bool, default true), ``strict`` (bool, default false; true: raise ``NotImplementedError`` instead of handing
anything to the reference).
"""
from __future__ import annotations

import importlib
import logging
import operator
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
    # matrix-level forms: DCT over a log-mel matrix, spectral contrast of a magnitude spectrogram.  Only the defining
    # modules are rebound: the reference manager's own imported copies serve its (reference-path) calls unchanged.
    ("sygnals.core.features.cepstral", "mfcc", ()),
    ("sygnals.core.features.frequency_domain", "spectral_contrast", ()),
]


def _length_ok(n, features: bool = False) -> bool:
    """Transform lengths the engine serves: powers of two in [32, 8192] on the register-FFT kernels; every other length >= 8 whose
    prime factors are at most 13 and whose buffers fit in one SM's shared memory on the mixed-radix kernels (syg_mixed.cuh;
    mirrors ``mixed_plan_of`` in csrc/syg_api.cu -- the library re-checks and the plugin falls back on NotImplementedError)."""
    try:
        n = operator.index(n)
    except TypeError:
        return False
    if n < 8:
        return False
    if 32 <= n <= 8192 and (n & (n - 1)) == 0:
        return True
    L = n // 2 if n % 2 == 0 else n
    m = L
    for p in (2, 3, 5, 7, 11, 13):
        while m % p == 0:
            m //= p
    if m != 1:
        return False
    B = n // 2 + 1
    words = 4 * L + ((B + (B >> 5) + 1 + 256) if features else 0) + 6 + 2 * (8 + 256)
    return words * 4 <= 227 * 1024


_LENGTH_RULE = "a power of two in [32, 8192] or a length >= 8 with prime factors <= 13 that fits in shared memory"


class SygnalsB200Plugin(SygnalsPluginBase):
    """Registers the engine's functions and rebinds the reference's hot-path names to them."""

    def __init__(self):
        self._saved: List[Tuple[Any, str, Any]] = []
        self._strict = False
        self._device: Optional[int] = None
        self._warned: set = set()

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
        if self._device is not None:
            import os
            os.environ["SYGB200_DEVICE"] = str(int(self._device))      # the rebound mirrors take their engine from _ffi.engine()
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
        self._note_reference(what, reason)
        return original

    def _note_reference(self, what: str, reason: str) -> None:
        """WARNING, once per distinct (call, reason): the user must be able to tell which implementation ran."""
        key = (what, reason)
        if key not in self._warned:
            self._warned.add(key)
            logger.warning(f"sygnals-b200: {what} runs on the REFERENCE (CPU) implementation: {reason}")

    def _call_engine_or_reference(self, what: str, engine_fn: Callable, original: Optional[Callable], *args, **kw):
        """The engine mirror; a NotImplementedError it raises later than the routing checks (n_mels > 256, entropy bins, ...)
        hands the call to the reference unless strict."""
        try:
            return engine_fn(*args, **kw)
        except NotImplementedError as e:
            if self._strict or original is None:
                raise
            self._note_reference(what, str(e))
            return original(*args, **kw)

    def make_extract_features(self, original: Optional[Callable]) -> Callable:
        from .core.features import manager as m

        def extract_features(y, sr, features, frame_length=2048, hop_length=512, center=True, window="hann",
                             feature_params=None, output_format="dataframe"):
            kw = dict(frame_length=frame_length, hop_length=hop_length, center=center, window=window, feature_params=feature_params,
                      output_format=output_format)
            reason = None
            names = sorted(m._ALL_KNOWN_FEATURES) if features == ["all"] else list(features)
            if not _length_ok(frame_length, features=True):
                reason = f"frame_length={frame_length} is not {_LENGTH_RULE}"
            elif not isinstance(window, str) or window.lower() not in _ffi.WINDOW_IDS:
                reason = f"window={window!r} is not built into the engine"
            if reason is not None:
                return self._engine_ok("extract_features", reason, original)(y, sr, features, **kw)
            missing = [f for f in names if f in m._ALL_KNOWN_FEATURES and f not in m.ENGINE_FEATURES]
            if not missing:
                return self._call_engine_or_reference("extract_features", m.extract_features, original, y, sr, features, **kw)
            ref = self._engine_ok("extract_features", f"no CUDA kernel for {missing}", original)
            have = [f for f in names if f not in missing]
            if not have:
                return ref(y, sr, features, **kw)
            # mixed request: kernelled features in one fused launch on the engine, the rest on the reference, merged in request order
            # (spectral_bandwidth before spectral_centroid injects the centroid column, manager.py:296-301: keep both on one side)
            kw_d = dict(kw, output_format="dict_of_arrays")
            got = self._call_engine_or_reference("extract_features", m.extract_features, original, y, sr, have, **kw_d)
            rest = ref(y, sr, missing, **kw_d)
            from .batch import feature_row_names
            merged = {"time": got.get("time", rest.get("time"))}
            seen = set()
            for f in names:
                if f in seen:
                    continue
                seen.add(f)
                src = rest if f in missing else got
                cols = [f] if f in missing else feature_row_names([f], feature_params)
                if f == "spectral_bandwidth" and "spectral_centroid" in src:
                    cols = ["spectral_centroid"] + cols              # the injected column (manager.py:296-301)
                for col in cols:
                    if col in src and col not in merged:
                        merged[col] = src[col]
            if output_format == "dict_of_arrays":
                return merged
            import pandas as pd
            idx = pd.to_timedelta(merged.pop("time"), unit="s")
            df = pd.DataFrame(merged, index=idx)
            df.index.name = "time"
            return df

        extract_features.__doc__ = m.extract_features.__doc__
        extract_features.__wrapped_reference__ = original
        return extract_features

    def make_compute_stft(self, original: Optional[Callable]) -> Callable:
        from .core import dsp as d

        def compute_stft(y, n_fft=2048, hop_length=None, win_length=None, window="hann", center=True, pad_mode="constant"):
            reason = None
            if not _length_ok(n_fft):
                reason = f"n_fft={n_fft} is not {_LENGTH_RULE}"
            elif not isinstance(window, str) or window.lower() not in _ffi.WINDOW_IDS:
                reason = f"window={window!r} is not built into the engine"
            elif pad_mode not in _ffi.PAD_IDS:
                reason = f"pad_mode={pad_mode!r} is not built into the engine"
            kw = dict(n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window, center=center, pad_mode=pad_mode)
            ref = self._engine_ok("compute_stft", reason, original)
            if ref is not None:
                return ref(y, **kw)
            return self._call_engine_or_reference("compute_stft", d.compute_stft, original, y, **kw)

        compute_stft.__wrapped_reference__ = original
        return compute_stft

    def _make_psd(self, which: str, original: Optional[Callable]) -> Callable:
        from .core import dsp as d
        mirror = getattr(d, which)

        pos_names = ("nperseg", "noverlap", "nfft", "detrend", "scaling") if which == "compute_psd_welch" else ("nfft", "detrend", "scaling")

        def psd(x, fs=1.0, window="hann", *args, **kw):
            if len(args) > len(pos_names):
                raise TypeError(f"{which}() takes at most {3 + len(pos_names)} positional arguments")
            for name, val in zip(pos_names, args):                   # the reference signature accepts these positionally
                if name in kw:
                    raise TypeError(f"{which}() got multiple values for argument '{name}'")
                kw[name] = val
            n = len(x) if hasattr(x, "__len__") else 0
            nfft = kw.get("nfft") or kw.get("nperseg") or (256 if which == "compute_psd_welch" else n)
            if which == "compute_psd_welch" and kw.get("nfft") is None:
                nfft = min(int(nfft), n) if n else nfft                 # scipy clamps nperseg to the signal length
            reason = None
            if not _length_ok(int(nfft)):
                reason = f"nfft={nfft} is not {_LENGTH_RULE}"
            elif not isinstance(window, str) or window.lower() not in _ffi.WINDOW_IDS:
                reason = f"window={window!r} is not built into the engine"
            elif kw.get("detrend", "constant") not in ("constant", True, False):
                reason = f"detrend={kw.get('detrend')!r} is not built into the engine"
            ref = self._engine_ok(which, reason, original)
            if ref is not None:
                return ref(x, fs=fs, window=window, **kw)
            return self._call_engine_or_reference(which, mirror, original, x, fs=fs, window=window, **kw)

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
            if not _length_ok(fl, features=True):
                reason = f"frame_length={fl} is not {_LENGTH_RULE}"
            elif kw.get("S") is not None:
                reason = "spectrogram input has no CUDA kernel"
            elif kw.get("pad_mode", "constant") != "constant":
                reason = f"pad_mode={kw.get('pad_mode')!r} is not built into the engine"
            elif set(kw) - {"S", "frame_length", "hop_length", "center", "pad_mode"}:
                reason = "extra librosa keyword arguments"
            ref = self._engine_ok(which, reason, original)
            if ref is not None:
                return ref(y, *args, **kw)
            return self._call_engine_or_reference(which, mirror, original, y, *args, **kw)

        feature.__name__ = which
        feature.__wrapped_reference__ = original
        return feature

    def make_mfcc(self, original: Optional[Callable]) -> Callable:
        from .core.features import cepstral as c

        def mfcc(y=None, sr=None, S=None, n_mfcc=13, dct_type=2, norm="ortho", lifter=0.0, **kwargs):
            reason = None
            if S is None:
                reason = "the time-series form mfcc(y=...) is librosa's own mel pipeline (no kernel; extract_features(['mfcc']) is the fused path)"
            elif kwargs:
                reason = f"librosa keyword arguments {sorted(kwargs)}"
            ref = self._engine_ok("mfcc", reason, original)
            kw = dict(y=y, sr=sr, S=S, n_mfcc=n_mfcc, dct_type=dct_type, norm=norm, lifter=lifter, **kwargs)
            if ref is not None:
                return ref(**kw)
            return self._call_engine_or_reference("mfcc", c.mfcc, original, **kw)

        mfcc.__wrapped_reference__ = original
        return mfcc

    def make_spectral_contrast(self, original: Optional[Callable]) -> Callable:
        from .core.features import frequency_domain as fd

        def spectral_contrast(S, sr, n_bands=6, fmin=200.0, freqs=None, **kwargs):
            reason = None
            if freqs is not None:
                reason = "custom bin frequencies"
            elif set(kwargs) - {"quantile"}:
                reason = f"librosa keyword arguments {sorted(kwargs)}"
            elif getattr(S, "ndim", 2) == 2 and S.shape[0] > 4097:
                reason = f"{S.shape[0]} frequency rows (the engine serves up to 4097)"
            ref = self._engine_ok("spectral_contrast", reason, original)
            if ref is not None:
                return ref(S, sr, n_bands=n_bands, fmin=fmin, freqs=freqs, **kwargs)
            return self._call_engine_or_reference("spectral_contrast", fd.spectral_contrast, original, S, sr, n_bands=n_bands, fmin=fmin,
                                                  freqs=freqs, **kwargs)

        spectral_contrast.__wrapped_reference__ = original
        return spectral_contrast

    def _replacement(self, attr: str, original: Optional[Callable]) -> Callable:
        if attr == "mfcc":
            return self.make_mfcc(original)
        if attr == "spectral_contrast":
            return self.make_spectral_contrast(original)
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
