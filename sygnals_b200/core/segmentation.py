"""Mirror of ``sygnals/core/segmentation.py:25-117`` (``segment_fixed_length``) for the B200 engine.

Segment boundaries come from the engine's C ABI (``syg_segment_table``: the reference's ``int()`` truncations and loop
conditions, segmentation.py:62-114).  The feature kernels never materialise segments (they frame by index, see
``sygnals_b200.batch.segment_features``); this function exists for drop-in callers that want the list of arrays.
"""
from __future__ import annotations

import logging
from typing import List, Optional

import numpy as np

from .. import _ffi

logger = logging.getLogger(__name__)


def segment_fixed_length(y, sr: int, segment_length_sec: float, overlap_ratio: float = 0.0, pad: bool = True,
                         min_segment_length_sec: Optional[float] = None) -> List[np.ndarray]:
    y = np.asarray(y)
    if y.ndim != 1:
        raise ValueError("Input signal y must be 1D.")
    if segment_length_sec <= 0:
        raise ValueError("segment_length_sec must be positive.")
    if not 0.0 <= overlap_ratio < 1.0:
        raise ValueError("overlap_ratio must be between 0.0 and < 1.0.")
    seg_len, seg_hop, starts, valid = _ffi.library().segment_table(len(y), sr, segment_length_sec, overlap_ratio, pad,
                                                                   min_segment_length_sec)
    if seg_len == 0:
        logger.warning(f"Segment length in samples is 0 for {segment_length_sec}s and sr={sr}. No segments generated.")
        return []
    out: List[np.ndarray] = []
    for s, v in zip(starts.tolist(), valid.tolist()):
        seg = y[s:s + v]
        if v < seg_len:                                   # only reachable with pad=True (segmentation.py:90-94)
            seg = np.pad(seg, (0, seg_len - v), mode="constant")
        out.append(seg.astype(np.float64, copy=False))
    logger.debug(f"Generated {len(out)} fixed-length segments.")
    return out


def segment_table(total_samples: int, sr: int, segment_length_sec: float, overlap_ratio: float = 0.0, pad: bool = True,
                  min_segment_length_sec: Optional[float] = None):
    """(seg_len, seg_hop, starts int64[n], valid int32[n]) -- the index form the kernels consume."""
    if segment_length_sec <= 0:
        raise ValueError("segment_length_sec must be positive.")
    if not 0.0 <= overlap_ratio < 1.0:
        raise ValueError("overlap_ratio must be between 0.0 and < 1.0.")
    return _ffi.library().segment_table(total_samples, sr, segment_length_sec, overlap_ratio, pad, min_segment_length_sec)
