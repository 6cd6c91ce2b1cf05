"""Mirror of ``sygnals/core/ml_utils/formatters.py:51-163`` (``format_feature_vectors_per_segment``) on the B200 engine: the
consumer of the feature matrix in ``sygnals save dataset`` (``sygnals/cli/save_cmd.py:140-190``).

Same signature, validation, warnings and output formats.  The NaN-aware mean / std / median / min / max over each segment's
frames (formatters.py:28-47) runs on the GPU (one warp per segment x feature, FP64 accumulation, exact median); the frame
features are taken as float32 (what the engine produces), results are float64.

:func:`aggregate_segments` is the device-resident form for the engine's own ``[n_segments, rows, T]`` output: aggregating before
the final gather shrinks the collective by a factor T.
"""
from __future__ import annotations

import logging
import warnings
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from ... import _ffi

logger = logging.getLogger(__name__)

AGGREGATIONS = tuple(_ffi.AGG_IDS)


def _agg_ids(feature_names: Sequence[str], aggregation) -> List[int]:
    if isinstance(aggregation, str):
        if aggregation not in _ffi.AGG_IDS:
            raise ValueError(f"Unknown global aggregation function: '{aggregation}'. Available: {list(_ffi.AGG_IDS.keys())}")
        return [_ffi.AGG_IDS[aggregation]] * len(feature_names)
    if isinstance(aggregation, dict):
        ids = []
        for name in feature_names:
            m = aggregation.get(name, "mean")
            if m not in _ffi.AGG_IDS:
                raise ValueError(f"Unknown aggregation function '{m}' for feature '{name}'. Available: {list(_ffi.AGG_IDS.keys())}")
            ids.append(_ffi.AGG_IDS[m])
        return ids
    raise TypeError("aggregation must be a string or a dictionary.")


def aggregate_segments(features, names: Sequence[str], aggregation: Union[str, Dict[str, str]] = "mean"):
    """``features``: CUDA float32 tensor ``[n_segments, rows, T]`` (``batch.segment_features(...)['features']``) ->
    CUDA float64 tensor ``[n_segments, rows]``; asynchronous on the current stream."""
    import torch
    if features.dim() != 3 or not features.is_cuda:
        raise ValueError("features must be a CUDA tensor [n_segments, rows, T]")
    f = features if (features.dtype == torch.float32 and features.is_contiguous()) else features.contiguous().float()
    n, rows, T = (int(x) for x in f.shape)
    ids = _agg_ids(list(names), aggregation)
    if len(ids) != rows:
        raise ValueError("names must have one entry per feature row")
    out = torch.full((n, rows), float("nan"), dtype=torch.float64, device=f.device)
    eng = _ffi.engine(f.device.index or 0)
    for r0 in range(0, rows, 64):                               # the kernel takes up to 64 rows per launch
        r1 = min(rows, r0 + 64)
        if n and T:
            part = out if (r0 == 0 and r1 == rows) else torch.empty((n, r1 - r0), dtype=torch.float64, device=f.device)
            import ctypes as C
            off = torch.arange(n, device=f.device, dtype=torch.int64) * (rows * T) + r0 * T
            eng.aggregate_dev(f.data_ptr(), n, r1 - r0, T, ids[r0:r1], part.data_ptr(), seg_off_ptr=off.data_ptr(), fixed_len=T,
                              stream=torch.cuda.current_stream(f.device).cuda_stream)
            del off                                            # stream-ordered allocator: safe to drop once the launch is queued
            if part is not out:
                out[:, r0:r1] = part
    return out


def format_feature_vectors_per_segment(features_dict: Dict[str, np.ndarray], segment_indices: List[Tuple[int, int]],
                                       aggregation: Union[str, Dict[str, str]] = "mean", output_format: str = "dataframe",
                                       segment_labels: Optional[List[Any]] = None):
    import pandas as pd
    if not features_dict:
        logger.warning("Input features_dict is empty. Returning empty result.")
        return pd.DataFrame() if output_format == "dataframe" else np.empty((0, 0), dtype=np.float64)
    feature_names = list(features_dict.keys())
    frame_counts = [len(arr) for arr in features_dict.values()]
    num_frames = frame_counts[0]
    if not all(c == num_frames for c in frame_counts):
        raise ValueError(f"All feature arrays in features_dict must have the same length. Found lengths: {frame_counts}")
    if not segment_indices:
        logger.warning("No segment indices provided. Returning empty result.")
        return pd.DataFrame() if output_format == "dataframe" else np.empty((0, len(feature_names)), dtype=np.float64)
    if segment_labels is not None and len(segment_labels) != len(segment_indices):
        raise ValueError("Length of segment_labels must match length of segment_indices.")
    ids = _agg_ids(feature_names, aggregation)
    if output_format not in ("dataframe", "numpy"):
        raise ValueError(f"Unknown output_format: '{output_format}'. Choose 'dataframe' or 'numpy'.")
    n_seg, n_feat = len(segment_indices), len(feature_names)
    off = np.zeros(n_seg, dtype=np.int64)
    ln = np.zeros(n_seg, dtype=np.int32)
    for i, (s, e) in enumerate(segment_indices):
        if not (0 <= s < num_frames and s < e and e <= num_frames):            # formatters.py:138-147: NaN row + warning
            msg = f"Invalid segment indices ({s}, {e}) for num_frames={num_frames}. Skipping segment {i}."
            logger.warning(msg)
            warnings.warn(msg, UserWarning, stacklevel=2)
            continue
        off[i], ln[i] = s, e - s
    import torch
    eng = _ffi.engine()
    dev = torch.device("cuda", eng.device)
    feats = torch.from_numpy(np.stack([np.asarray(features_dict[k], dtype=np.float32) for k in feature_names])).to(dev)   # [rows, N]
    out = np.full((n_seg, n_feat), np.nan, dtype=np.float64)
    d_off, d_len = torch.from_numpy(off).to(dev), torch.from_numpy(ln).to(dev)
    for r0 in range(0, n_feat, 64):
        r1 = min(n_feat, r0 + 64)
        part = torch.empty((n_seg, r1 - r0), dtype=torch.float64, device=dev)
        eng.aggregate_dev(feats[r0:r1].data_ptr(), n_seg, r1 - r0, num_frames, ids[r0:r1], part.data_ptr(), seg_off_ptr=d_off.data_ptr(),
                          seg_len_ptr=d_len.data_ptr(), stream=torch.cuda.current_stream(dev).cuda_stream)
        out[:, r0:r1] = part.cpu().numpy()
    if output_format == "dataframe":
        index = segment_labels if segment_labels is not None else pd.RangeIndex(n_seg, name="segment_index")
        return pd.DataFrame(out, columns=feature_names, index=index)
    return out
