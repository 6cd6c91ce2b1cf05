"""Mirror of ``sygnals/core/features/frequency_domain.py:147-212`` (``spectral_contrast``) on the B200 engine.

``spectral_contrast(S, sr, n_bands, fmin)`` takes a magnitude spectrogram (frequency x time) and returns ``(n_bands + 1, T)``
float64, exactly as the reference / librosa do: octave bands from ``fmin``, per band the mean of the top and of the bottom
``quantile`` of the bins, ``power_to_db(peak) - power_to_db(valley)`` with each array clamped 80 dB below its own maximum.  The
selection runs in the same warp-level code as the fused feature kernel (``syg_spectral_contrast_from_mag_f32``), in FP32.

Not served (``NotImplementedError``; the plugin hands these to the reference): custom ``freqs``, ``linear=True`` and other librosa
keyword arguments.  The per-frame scalar helpers of the same module (``spectral_centroid(spectrum, freqs)``, ``spectral_rolloff``,
...) are the bodies of the manager's Python loops; the engine replaces the loops (``extract_features``), the helpers themselves
stay the reference's -- one launch per 1025-bin spectrum would lose to numpy.
"""
from __future__ import annotations

import logging
from typing import Any, Optional

import numpy as np

from ... import _ffi

logger = logging.getLogger(__name__)


def spectral_contrast(S: np.ndarray, sr: int, n_bands: int = 6, fmin: float = 200.0, freqs: Optional[np.ndarray] = None,
                      **kwargs: Any) -> np.ndarray:
    S = np.asarray(S)
    if S.ndim != 2:
        raise ValueError("Input S must be a 2D spectrogram (frequency x time).")
    if freqs is not None:
        raise NotImplementedError("spectral_contrast(freqs=...): custom bin frequencies are not built into the B200 engine")
    quantile = float(kwargs.pop("quantile", 0.02))
    if kwargs:
        raise NotImplementedError(f"spectral_contrast: keyword arguments {sorted(kwargs)} are not built into the B200 engine")
    if np.any(S < 0):
        logger.warning("Input spectrogram S contains negative values. Using absolute values.")
        S = np.abs(S)
    logger.debug(f"Calculating Spectral Contrast: n_bands={n_bands}, fmin={fmin}")
    import torch
    B, T = int(S.shape[0]), int(S.shape[1])
    eng = _ffi.engine()
    dev = torch.device("cuda", eng.device)
    d_S = torch.from_numpy(np.ascontiguousarray(S, dtype=np.float32)).to(dev)
    d_out = torch.empty((int(n_bands) + 1, T), dtype=torch.float32, device=dev)
    eng.spectral_contrast_from_mag_dev(d_S.data_ptr(), B, T, float(sr), int(n_bands), float(fmin), quantile, d_out.data_ptr(),
                                       torch.cuda.current_stream(dev).cuda_stream)
    return d_out.cpu().numpy().astype(np.float64)
