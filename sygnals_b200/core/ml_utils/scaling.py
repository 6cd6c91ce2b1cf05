"""Sharded counterpart of ``sygnals/core/ml_utils/scaling.py:49-175`` (``apply_scaling`` with the ``'standard'`` and ``'minmax'``
scalers of scikit-learn) for the feature matrix the engine produces: every rank holds a block of rows ``[n_local, n_features]``
(``dist.segment_features_sharded(..., aggregation=...)`` without the final gather), the column statistics are one
``all_reduce`` of (count, sum, sum of squares) or (min, max), and the transform is applied where the rows live.

The fitted object carries scikit-learn's attribute names (``mean_``, ``var_``, ``scale_``, ``n_samples_seen_`` /
``data_min_``, ``data_max_``, ``data_range_``, ``scale_``, ``min_``) so it can stand in for ``apply_scaling``'s second return
value; NaNs are ignored in ``fit`` and kept in ``transform`` exactly as scikit-learn does.  torch tensors (CUDA -> NCCL, CPU ->
gloo) or numpy arrays.  ``'robust'`` needs order statistics of whole columns: the (small) row blocks are all-gathered once
and every rank takes the NaN-aware median / quantiles of the full columns with numpy's linear interpolation rule.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np


@dataclass
class FittedScaler:
    kind: str
    n_samples_seen_: np.ndarray
    scale_: np.ndarray
    mean_: Optional[np.ndarray] = None          # standard
    var_: Optional[np.ndarray] = None
    with_mean: bool = True
    with_std: bool = True
    data_min_: Optional[np.ndarray] = None      # minmax
    data_max_: Optional[np.ndarray] = None
    data_range_: Optional[np.ndarray] = None
    min_: Optional[np.ndarray] = None
    feature_range: Tuple[float, float] = (0.0, 1.0)
    center_: Optional[np.ndarray] = None        # robust
    quantile_range: Tuple[float, float] = (25.0, 75.0)
    with_centering: bool = True
    with_scaling: bool = True
    _dev: dict = field(default_factory=dict, repr=False)

    def transform(self, X):
        is_np = isinstance(X, np.ndarray)
        if is_np:
            if self.kind == "standard":
                out = X.astype(np.float64, copy=True)
                if self.with_mean:
                    out -= self.mean_
                if self.with_std:
                    out /= self.scale_
                return out
            if self.kind == "robust":
                out = X.astype(np.float64, copy=True)
                if self.with_centering:
                    out -= self.center_
                if self.with_scaling:
                    out /= self.scale_
                return out
            return X.astype(np.float64) * self.scale_ + self.min_
        import torch
        key = (X.device, X.dtype)
        if key not in self._dev:
            t = lambda a: None if a is None else torch.as_tensor(a, dtype=torch.float64, device=X.device)  # noqa: E731
            self._dev[key] = (t(self.mean_ if self.kind != "robust" else self.center_), t(self.scale_), t(self.min_))
        mean, scale, mn = self._dev[key]
        Xd = X.to(torch.float64)
        if self.kind == "standard":
            if self.with_mean:
                Xd = Xd - mean
            return Xd / scale if self.with_std else Xd
        if self.kind == "robust":
            if self.with_centering:
                Xd = Xd - mean
            return Xd / scale if self.with_scaling else Xd
        return Xd * scale + mn


def _handle_zeros(scale: np.ndarray) -> np.ndarray:
    """sklearn.preprocessing._data._handle_zeros_in_scale: (near-)constant columns are left unscaled."""
    s = scale.copy()
    s[s < 10 * np.finfo(np.float64).eps] = 1.0
    return s


def _nan_quantiles(X, qs):
    """Per-column quantiles ignoring NaNs, numpy's default ('linear') rule: h = (n - 1) q, v[floor h] + (h - floor h) (v[ceil h] - v[floor h]).
    X: float64 [rows, cols] torch tensor; returns [len(qs), cols] (NaN for an all-NaN column)."""
    import torch
    srt, _ = torch.sort(X, dim=0)                                  # NaNs sort last
    n = (~torch.isnan(X)).sum(0)                                   # valid rows per column
    out = []
    rows = X.shape[0]
    for q in qs:
        h = (n.to(torch.float64) - 1.0) * float(q)
        lo = torch.clamp(torch.floor(h).to(torch.int64), 0, max(rows - 1, 0))
        hi = torch.clamp(torch.ceil(h).to(torch.int64), 0, max(rows - 1, 0))
        vlo = torch.gather(srt, 0, lo[None, :])[0]
        vhi = torch.gather(srt, 0, hi[None, :])[0]
        g = h - torch.floor(h)
        # numpy's _lerp: the upper form for g >= 0.5 (identical rounding to np.percentile / np.nanmedian)
        d = vhi - vlo
        v = torch.where(g >= 0.5, vhi - d * (1.0 - g), vlo + d * g)
        out.append(torch.where(n > 0, v, torch.full_like(v, float("nan"))))
    return torch.stack(out)


def fit_scaler(X_local, scaler_type: str = "standard", with_mean: bool = True, with_std: bool = True,
               feature_range: Tuple[float, float] = (0.0, 1.0), with_centering: bool = True, with_scaling: bool = True,
               quantile_range: Tuple[float, float] = (25.0, 75.0), group=None) -> FittedScaler:
    """Column statistics over ALL ranks' rows (one all-reduce when a process group is initialised)."""
    import torch
    import torch.distributed as dist
    if scaler_type not in ("standard", "minmax", "robust"):
        raise ValueError(f"Unsupported scaler_type: '{scaler_type}'. Supported types: ['standard', 'minmax', 'robust']")
    X = torch.as_tensor(X_local)
    if X.dim() == 1:
        X = X.reshape(-1, 1)
    if X.dim() != 2:
        raise ValueError(f"Input features must be 1D or 2D (samples/frames x features), got shape {tuple(X.shape)}")
    X = X.to(torch.float64)
    nan = torch.isnan(X)
    world = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if scaler_type == "robust":
        # RobustScaler (scaling.py:106-111 -> sklearn): center_ = nanmedian, scale_ = nanpercentile(q_max) - nanpercentile(q_min)
        q_min, q_max = quantile_range
        if not 0 <= q_min <= q_max <= 100:
            raise ValueError(f"Invalid quantile range: {quantile_range}")
        full = X
        if world:                                                   # order statistics need the whole column: gather the row blocks
            cnt = torch.tensor([X.shape[0]], dtype=torch.int64, device=X.device)
            cnts = [torch.zeros_like(cnt) for _ in range(dist.get_world_size(group))]
            dist.all_gather(cnts, cnt, group=group)
            mx = int(max(int(c.item()) for c in cnts))
            pad = torch.full((mx, X.shape[1]), float("nan"), dtype=torch.float64, device=X.device)
            pad[: X.shape[0]] = X
            parts = [torch.empty_like(pad) for _ in cnts]
            dist.all_gather(parts, pad, group=group)
            full = torch.cat([p_[: int(c.item())] for p_, c in zip(parts, cnts)], dim=0)
        qv = _nan_quantiles(full, [0.5, q_min / 100.0, q_max / 100.0]).cpu().numpy()
        n_seen = (~torch.isnan(full)).sum(0).cpu().numpy().astype(np.int64)
        scale = _handle_zeros(qv[2] - qv[1]) if with_scaling else np.ones(X.shape[1])
        return FittedScaler("robust", n_seen, scale, center_=qv[0] if with_centering else None, quantile_range=(q_min, q_max),
                            with_centering=with_centering, with_scaling=with_scaling)
    if scaler_type == "standard":
        Z = torch.where(nan, torch.zeros_like(X), X)
        stats = torch.stack([(~nan).sum(0).to(torch.float64), Z.sum(0)])
        if world:
            dist.all_reduce(stats, group=group)
        n = stats[0]
        mean = stats[1] / n
        # second pass around the global mean (no cancellation), second all-reduce of one row
        ss = torch.where(nan, torch.zeros_like(X), (X - mean) ** 2).sum(0)
        if world:
            dist.all_reduce(ss, group=group)
        var = (ss / n).cpu().numpy()
        scale = _handle_zeros(np.sqrt(var)) if with_std else np.ones_like(var)
        return FittedScaler("standard", n.cpu().numpy().astype(np.int64), scale, mean_=mean.cpu().numpy(), var_=var,
                            with_mean=with_mean, with_std=with_std)
    big = torch.finfo(torch.float64).max
    mn = torch.where(nan, torch.full_like(X, big), X).amin(0) if X.shape[0] else torch.full((X.shape[1],), big, dtype=torch.float64, device=X.device)
    mx = torch.where(nan, torch.full_like(X, -big), X).amax(0) if X.shape[0] else torch.full((X.shape[1],), -big, dtype=torch.float64, device=X.device)
    cnt = (~nan).sum(0).to(torch.float64)
    if world:
        dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(cnt, group=group)
    dmin, dmax = mn.cpu().numpy(), mx.cpu().numpy()
    rng = dmax - dmin
    lo, hi = feature_range
    scale = (hi - lo) / _handle_zeros(rng)
    return FittedScaler("minmax", cnt.cpu().numpy().astype(np.int64), scale, data_min_=dmin, data_max_=dmax, data_range_=rng,
                        min_=lo - dmin * scale, feature_range=(lo, hi))


def apply_scaling(features, scaler_type: str = "standard", scaler_params: Optional[dict] = None, fit: bool = True,
                  scaler_instance: Optional[FittedScaler] = None, group=None):
    """Mirror of ``apply_scaling`` (scaling.py:49-142) on this rank's block of rows: (scaled block, fitted scaler)."""
    if fit:
        sc = fit_scaler(features, scaler_type, group=group, **(scaler_params or {}))
    else:
        if scaler_instance is None:
            raise ValueError("`scaler_instance` must be provided when `fit=False`.")
        sc = scaler_instance
    X = features.reshape(-1, 1) if getattr(features, "ndim", 2) == 1 else features
    return sc.transform(X), sc
