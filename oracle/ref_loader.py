"""
TEST INFRASTRUCTURE ONLY (this container only -- /root/reference does not exist
on the GPU box).

Imports the UNMODIFIED reference package from ``/root/reference`` with the six
runtime dependencies that are not installed offline replaced by stubs:
``librosa`` -> ``oracle.librosa_shim``; ``soundfile``, ``pandasql``, ``pywt``,
``matplotlib``(+pyplot), ``resampy`` -> inert modules (never called on the hot
path).  Used by ``tests/golden/make_golden.py`` to generate golden vectors and by
``tests/test_oracle_vs_reference.py`` to validate ``oracle.sygnals_oracle``.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SYGNALS_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "sygnals", "core"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def install_stubs() -> None:
    from . import librosa_shim

    if "librosa" not in sys.modules or getattr(sys.modules["librosa"], "__version__", "") .endswith("-shim"):
        sys.modules["librosa"] = librosa_shim
        for sub in ("feature", "util", "effects", "onset", "filters", "core"):
            sys.modules[f"librosa.{sub}"] = getattr(librosa_shim, sub)
        sys.modules["librosa.util.exceptions"] = librosa_shim.util.exceptions

    def _missing(*a, **k):
        raise NotImplementedError("stubbed dependency (not on the hot path)")

    if "soundfile" not in sys.modules:
        sys.modules["soundfile"] = _stub(
            "soundfile", available_formats=lambda: {"WAV": "WAV", "FLAC": "FLAC", "OGG": "OGG"},
            available_subtypes=lambda fmt=None: {"PCM_16": "", "PCM_24": "", "FLOAT": ""},
            read=_missing, write=_missing, info=_missing, SoundFile=object,
            LibsndfileError=RuntimeError, SoundFileError=RuntimeError)
    if "pandasql" not in sys.modules:
        sys.modules["pandasql"] = _stub("pandasql", sqldf=_missing)
    if "pywt" not in sys.modules:
        sys.modules["pywt"] = _stub("pywt", wavedec=_missing, waverec=_missing, wavelist=lambda *a, **k: [],
                                    Wavelet=object, dwt_max_level=_missing)
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib", use=lambda *a, **k: None)
        plt = _stub("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
        for sub in ("figure", "axes", "colors", "cm", "ticker"):
            sm = _stub(f"matplotlib.{sub}", Figure=object, Axes=object)
            setattr(mpl, sub, sm)
            sys.modules[f"matplotlib.{sub}"] = sm
    if "resampy" not in sys.modules:
        sys.modules["resampy"] = _stub("resampy", resample=_missing)


def load_reference():
    """Return the unmodified reference ``sygnals`` package (import side effects
    limited to sys.modules / sys.path)."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import sygnals  # noqa: F401
    import sygnals.core.dsp  # noqa: F401
    import sygnals.core.segmentation  # noqa: F401
    import sygnals.core.features.manager  # noqa: F401
    import sygnals.core.features.cepstral  # noqa: F401
    import sygnals.core.features.frequency_domain  # noqa: F401
    import sygnals.core.features.time_domain  # noqa: F401
    import sygnals.core.audio.features  # noqa: F401
    return sys.modules["sygnals"]
