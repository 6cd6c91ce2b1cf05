"""Mirror of ``sygnals/core/features/cepstral.py:20-120`` (``mfcc``) on the B200 engine, spectrogram form.

``mfcc(S=log_mel)`` -- the DCT over a caller-supplied log-power mel matrix, ``scipy.fftpack.dct(S, axis=-2, type, norm)[:n_mfcc]``
plus the sinusoidal lifter -- runs on the GPU in float64 (``syg_mfcc_from_logmel_f64``).  The time-series form ``mfcc(y=..., sr=...)``
is librosa's own mel pipeline (``power_to_db`` against ``ref=1.0``, not the manager's ``ref=np.max``); it has no kernel here and
raises ``NotImplementedError`` -- the plugin hands it to the reference.  Same signature, checks and messages as the reference.
"""
from __future__ import annotations

import logging
from typing import Any, Optional

import numpy as np

from ... import _ffi

logger = logging.getLogger(__name__)


def mfcc(y: Optional[np.ndarray] = None, sr: Optional[int] = None, S: Optional[np.ndarray] = None, n_mfcc: int = 13, dct_type: int = 2,
         norm: Optional[str] = "ortho", lifter: float = 0.0, **kwargs: Any) -> np.ndarray:
    logger.debug(f"Calculating MFCCs: n_mfcc={n_mfcc}, dct_type={dct_type}, norm={norm}, lifter={lifter}, kwargs={kwargs}")
    if S is None and y is None:
        raise ValueError("Either audio time series 'y' or Mel spectrogram 'S' must be provided.")
    if S is None and sr is None:
        raise ValueError("Sampling rate 'sr' must be provided when calculating MFCCs from time series 'y'.")
    if S is None:
        raise NotImplementedError("mfcc(y=...): the time-series form runs librosa's own mel pipeline and has no CUDA kernel in sygnals_b200 "
                                  "(use extract_features(['mfcc']) for the engine's fused path)")
    if y is not None:
        logger.warning("Both 'y' and 'S' provided for MFCC calculation. Using pre-computed 'S'. "
                       "Ensure 'S' is a log-power Mel spectrogram for correct results.")
    if norm not in ("ortho", None):
        raise ValueError(f"norm={norm!r} is not a valid DCT normalisation")
    if lifter < 0:
        raise ValueError(f"MFCC lifter={lifter} must be a non-negative number")
    import torch
    S = np.asarray(S, dtype=np.float64)
    if S.ndim < 2:
        raise ValueError("S must have at least two dimensions (n_mels, n_frames)")
    lead = S.shape[:-2]
    n_mels, T = int(S.shape[-2]), int(S.shape[-1])
    n_units = int(np.prod(lead)) if lead else 1
    C = min(int(n_mfcc), n_mels)
    eng = _ffi.engine()
    dev = torch.device("cuda", eng.device)
    d_S = torch.from_numpy(np.ascontiguousarray(S).reshape(n_units, n_mels, T)).to(dev)
    d_out = torch.empty((n_units, C, T), dtype=torch.float64, device=dev)
    eng.mfcc_from_logmel_dev(d_S.data_ptr(), n_units, n_mels, T, int(n_mfcc), int(dct_type), norm == "ortho", float(lifter), d_out.data_ptr(),
                             torch.cuda.current_stream(dev).cuda_stream)
    return d_out.cpu().numpy().reshape(lead + (C, T))


CEPSTRAL_FEATURES = {"mfcc": mfcc}
