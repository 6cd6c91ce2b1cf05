"""Mirror of the two array-level frame features of ``sygnals/core/audio/features.py`` that have a CUDA kernel:
``zero_crossing_rate`` (:26-71, librosa's edge padding when centred) and ``rms_energy`` (:73-131, librosa computes it in float32 on a
zero-padded signal).  Same signatures and return types (1-D float64, one value per frame).  They run the same fused kernel as
``extract_features``; the spectrogram form ``rms_energy(S=...)`` and extra librosa keyword arguments have no kernel and raise
``NotImplementedError`` (the plugin routes those calls to the reference)."""
from __future__ import annotations

import logging
from typing import Any, Optional

import numpy as np

from ..features import manager

logger = logging.getLogger(__name__)


def _frame_feature(name: str, y, frame_length: int, hop_length: int, center: bool) -> np.ndarray:
    y = np.asarray(y)
    if y.ndim != 1:
        raise ValueError("Input audio data must be a 1D array." if name == "zero_crossing_rate"
                         else "Input audio data 'y' must be a 1D array.")
    r = manager.extract_features(y, 1, [name], frame_length=frame_length, hop_length=hop_length, center=center,
                                 output_format="dict_of_arrays")          # sr is irrelevant for time-domain features
    return r.get(name, np.zeros(0, dtype=np.float64)).astype(np.float64, copy=False)


def zero_crossing_rate(y, frame_length: int = 2048, hop_length: int = 512, center: bool = True, **kwargs: Any) -> np.ndarray:
    if kwargs:
        raise NotImplementedError(f"zero_crossing_rate: keyword arguments {sorted(kwargs)} are not built into the B200 engine")
    logger.debug(f"Calculating Zero Crossing Rate: frame={frame_length}, hop={hop_length}, center={center}")
    return _frame_feature("zero_crossing_rate", y, frame_length, hop_length, center)


def rms_energy(y=None, *, S: Optional[np.ndarray] = None, frame_length: int = 2048, hop_length: int = 512, center: bool = True,
               pad_mode: str = "constant", **kwargs: Any) -> np.ndarray:
    if S is None and y is None:
        raise ValueError("Either audio time series 'y' or magnitude spectrogram 'S' must be provided.")
    if S is not None:
        raise NotImplementedError("rms_energy(S=...): the spectrogram form has no CUDA kernel in sygnals_b200")
    if pad_mode != "constant" and center:
        raise NotImplementedError(f"rms_energy: pad_mode={pad_mode!r} is not built into the B200 engine (constant only)")
    if kwargs:
        raise NotImplementedError(f"rms_energy: keyword arguments {sorted(kwargs)} are not built into the B200 engine")
    logger.debug(f"Calculating RMS Energy: frame={frame_length}, hop={hop_length}, center={center}")
    return _frame_feature("rms_energy", y, frame_length, hop_length, center)
