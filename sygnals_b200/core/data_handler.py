"""Egress of the B200 engine: feature blocks that live in HBM written in the formats ``sygnals.core.data_handler.save_data``
produces (``sygnals/core/data_handler.py:245-275``): an ``.npz`` with one array per feature name (``np.savez(**{col: values})``,
the layout ``sygnals features extract -o feats.npz`` writes and ``sygnals save dataset`` reads back, data_handler.py:160-169) or
a headerless / headered ``.csv``.

The device tensor is copied to page-locked host memory in one transfer (float64 on the way, as the reference's columns are) and
handed to numpy / pandas; file formats and key names are the reference's, so the reference's ``read_data`` loads the result.
"""
from __future__ import annotations

import os
from typing import Optional, Sequence, Union

import numpy as np


def _to_host(x) -> np.ndarray:
    if type(x).__module__.startswith("torch"):
        import torch
        if x.is_cuda:
            host = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
            host.copy_(x, non_blocking=True)
            torch.cuda.current_stream(x.device).synchronize()
            return host.numpy()
        return x.numpy()
    return np.asarray(x)


def save_feature_rows(path: Union[str, os.PathLike], names: Sequence[str], rows, times: Optional[np.ndarray] = None) -> None:
    """``rows``: ``[n_rows, T]`` (one unit of ``extract_features_batch`` / ``segment_features``; CUDA tensor or numpy).
    ``.npz``: ``{'time': times, name: float64[T], ...}`` exactly as ``save_data(extract_features(..., 'dataframe'))`` stores the frame
    features (data_handler.py:253-255); ``.csv``: one column per feature with a header row (``DataFrame.to_csv(index=False)``)."""
    path = os.fspath(path)
    r = _to_host(rows).astype(np.float64)
    if r.ndim != 2 or r.shape[0] != len(names):
        raise ValueError(f"rows must be [len(names), T], got {r.shape} for {len(names)} names")
    cols = {}
    if times is not None:
        cols["time"] = np.asarray(times, dtype=np.float64)
    cols.update({n: np.ascontiguousarray(r[i]) for i, n in enumerate(names)})
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npz":
        np.savez(path, **cols)
    elif ext == ".csv":
        import pandas as pd
        pd.DataFrame(cols).to_csv(path, index=False)
    else:
        raise ValueError(f"Unsupported output file format: '{ext}'. Supported formats: {{'.npz', '.csv'}}")


def save_segment_vectors(path: Union[str, os.PathLike], names: Sequence[str], vectors, labels: Optional[Sequence] = None) -> None:
    """``vectors``: ``[n_segments, n_rows]`` float64 (``dist.run_shard(..., aggregation=...)`` / ``syg_segment_vectors_*``), the
    matrix ``sygnals save dataset --assembly-method vectors`` assembles (save_cmd.py:140-190) and hands to ``save_data``: a 2-D numpy
    array goes to ``.npz`` under the key ``data`` (data_handler.py:257-259; ``feature_names`` / ``labels`` are additive keys), a
    DataFrame with one column per feature to ``.csv`` (``to_csv(index=False)``, data_handler.py:250-251)."""
    path = os.fspath(path)
    v = _to_host(vectors).astype(np.float64)
    if v.ndim != 2 or v.shape[1] != len(names):
        raise ValueError(f"vectors must be [n_segments, len(names)], got {v.shape} for {len(names)} names")
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npz":
        out = {"data": v, "feature_names": np.asarray(list(names))}
        if labels is not None:
            out["labels"] = np.asarray(labels)
        np.savez(path, **out)
    elif ext == ".csv":
        import pandas as pd
        pd.DataFrame(v, columns=list(names)).to_csv(path, index=False)
    else:
        raise ValueError(f"Unsupported output file format: '{ext}'. Supported formats: {{'.npz', '.csv'}}")
