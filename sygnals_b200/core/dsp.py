"""Mirror of the hot-path functions of ``sygnals/core/dsp.py`` on the B200 engine:

* ``compute_stft``             dsp.py:167-229  (librosa.stft semantics: periodic window zero-padded to n_fft and centred,
                               centre padding ``n_fft//2`` with ``pad_mode``, hop default ``win_length//4``, no
                               normalisation, ``(1 + n_fft/2, T)`` complex128)
* ``compute_psd_welch``        dsp.py:495-560  (scipy.signal.welch, one-sided)
* ``compute_psd_periodogram``  dsp.py:434-493  (scipy.signal.periodogram, one-sided)

The arithmetic runs in FP32 on the GPU (libsygb200.so); results are widened to the reference's dtypes.
"""
from __future__ import annotations

import logging
import warnings
from typing import Optional, Tuple, Union

import numpy as np

from .. import _ffi

logger = logging.getLogger(__name__)


def _window_id(window) -> int:
    if not isinstance(window, str) or window.lower() not in _ffi.WINDOW_IDS:
        raise NotImplementedError(f"window={window!r}: the B200 engine implements {sorted(set(_ffi.WINDOW_IDS))}")
    return _ffi.WINDOW_IDS[window.lower()]


def compute_stft(y, n_fft: int = 2048, hop_length: Optional[int] = None, win_length: Optional[int] = None,
                 window: str = "hann", center: bool = True, pad_mode: str = "constant") -> np.ndarray:
    y = np.asarray(y)
    if y.ndim != 1:
        raise ValueError("Input data must be a 1D array.")
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = int(win_length // 4)
    if hop_length <= 0:
        raise ValueError(f"hop_length={hop_length} must be a positive integer")
    if pad_mode not in _ffi.PAD_IDS:
        raise NotImplementedError(f"pad_mode={pad_mode!r}: the B200 engine implements {sorted(_ffi.PAD_IDS)}")
    n = len(y)
    if center:
        if n_fft > n:
            warnings.warn(f"n_fft={n_fft} is too large for input signal of length={n}")
    elif n < n_fft:
        raise ValueError(f"n_fft={n_fft} is too large for uncentered analysis of input signal of length={n}")
    eng = _ffi.engine()
    D = eng.stft_host(y.astype(np.float32, copy=False), eng.units_clips(1, n), n_fft, hop_length, win_length,
                      _window_id(window), center, _ffi.PAD_IDS[pad_mode], _ffi.OUT_COMPLEX)
    return D[0].astype(np.complex128)


def _welch_common(x, fs, window, nperseg, noverlap, nfft, detrend, scaling) -> Tuple[np.ndarray, np.ndarray]:
    n = len(x)
    if scaling not in _ffi.SCALING_IDS:
        raise ValueError(f"Unknown scaling: {scaling!r}")
    if detrend in ("constant", True):
        det = True
    elif detrend is False:
        det = False
    else:
        raise NotImplementedError(f"detrend={detrend!r}: the B200 engine implements 'constant' and False")
    if nperseg > n:
        warnings.warn(f"nperseg = {nperseg} is greater than input length  = {n}, using nperseg = {n}")
        nperseg = n
    if nfft is None:
        nfft = nperseg
    elif nfft < nperseg:
        raise ValueError("nfft must be greater than or equal to nperseg.")
    if noverlap is None:
        noverlap = nperseg // 2
    if noverlap >= nperseg:
        raise ValueError("noverlap must be less than nperseg.")
    eng = _ffi.engine()
    psd = eng.psd_welch_host(np.asarray(x, dtype=np.float32), eng.units_clips(1, n), float(fs), _window_id(window), int(nperseg),
                             int(noverlap), int(nfft), det, _ffi.SCALING_IDS[scaling])
    freqs = np.fft.rfftfreq(int(nfft), 1.0 / fs)
    return freqs.astype(np.float64, copy=False), psd[0].astype(np.float64)


def compute_psd_welch(x, fs: float = 1.0, window: str = "hann", nperseg: Optional[int] = None, noverlap: Optional[int] = None,
                      nfft: Optional[int] = None, detrend: Union[str, bool] = "constant",
                      scaling: str = "density") -> Tuple[np.ndarray, np.ndarray]:
    x = np.asarray(x)
    if x.ndim != 1:
        raise ValueError("Input data must be a 1D array.")
    if nperseg is None:
        nperseg = 256                                    # scipy.signal.welch default
    return _welch_common(x, fs, window, int(nperseg), noverlap, nfft, detrend, scaling)


def compute_psd_periodogram(x, fs: float = 1.0, window: str = "hann", nfft: Optional[int] = None,
                            detrend: Union[str, bool] = "constant", scaling: str = "density") -> Tuple[np.ndarray, np.ndarray]:
    x = np.asarray(x)
    if x.ndim != 1:
        raise ValueError("Input data must be a 1D array.")
    n = len(x)
    if nfft is not None and nfft < n:                     # scipy truncates the signal to nfft samples
        x, n = x[:nfft], int(nfft)
    return _welch_common(x, fs, window, n, 0, nfft, detrend, scaling)
