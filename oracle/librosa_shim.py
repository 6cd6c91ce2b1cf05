"""
TEST INFRASTRUCTURE ONLY -- numpy/scipy restatement of the librosa (>=0.10,
reference developed on 0.11.0; /root/reference/pyproject.toml:38) entry points
that the Sygnals hot path calls.  librosa's source is not under /root/reference
and cannot be installed offline, so each function restates librosa's published
algorithm.  The reference call site that depends on each one is cited.

Usable two ways:
  * imported as ``oracle.librosa_shim`` by ``oracle.sygnals_oracle``;
  * installed as ``sys.modules['librosa']`` by ``oracle.ref_loader`` so the
    UNMODIFIED reference runs on top of it (this container only).
"""
from __future__ import annotations

import types
import numpy as np
import scipy.fftpack
import scipy.signal

__version__ = "0.11.0-shim"


class ParameterError(Exception):
    """librosa.util.exceptions.ParameterError stand-in."""


# ----------------------------------------------------------------------------
# util
# ----------------------------------------------------------------------------
def _pad_center(data, *, size, axis=-1, **kwargs):
    """librosa.util.pad_center: centre ``data`` in a length-``size`` array."""
    kwargs.setdefault("mode", "constant")
    n = data.shape[axis]
    lpad = int((size - n) // 2)
    lengths = [(0, 0)] * data.ndim
    lengths[axis] = (lpad, int(size - n - lpad))
    if lpad < 0:
        raise ParameterError(f"Target size ({size}) must be at least input size ({n})")
    return np.pad(data, lengths, **kwargs)


def _frame(x, *, frame_length, hop_length, axis=-1, writeable=False, subok=False):
    """librosa.util.frame -- (..., frame_length, n_frames) strided view.

    Used by sygnals/core/features/manager.py:271-273 and inside rms / zcr.
    """
    x = np.asarray(x)
    if x.shape[axis] < frame_length:
        raise ParameterError(
            f"Input is too short (n={x.shape[axis]:d}) for frame_length={frame_length:d}"
        )
    if hop_length < 1:
        raise ParameterError(f"Invalid hop_length: {hop_length:d}")
    xw = np.lib.stride_tricks.sliding_window_view(x, frame_length, axis=axis)
    # sliding_window_view appends the window axis last; librosa moves it so the
    # result is (..., frame_length, n_frames) for axis=-1.
    if axis < 0:
        target_axis = axis - 1
    else:
        target_axis = axis + 1
    xw = np.moveaxis(xw, -1, target_axis)
    slices = [slice(None)] * xw.ndim
    slices[axis] = slice(0, None, hop_length)
    return xw[tuple(slices)]


def _abs2(x, dtype=None):
    """librosa.util.abs2"""
    if np.iscomplexobj(x):
        y = x.real ** 2 + x.imag ** 2
        return y if dtype is None else y.astype(dtype)
    return np.square(x, dtype=dtype)


def _expand_to(x, *, ndim, axes):
    axes_tup = (axes,) if isinstance(axes, int) else tuple(axes)
    shape = [1] * ndim
    for i, axi in enumerate(axes_tup):
        shape[axi] = x.shape[i]
    return x.reshape(shape)


util = types.ModuleType("librosa.util")
util.pad_center = _pad_center
util.frame = _frame
util.abs2 = _abs2
util.expand_to = _expand_to
util.exceptions = types.ModuleType("librosa.util.exceptions")
util.exceptions.ParameterError = ParameterError


# ----------------------------------------------------------------------------
# time / frequency conversions
# ----------------------------------------------------------------------------
def fft_frequencies(*, sr=22050, n_fft=2048):
    """librosa.fft_frequencies (manager.py:199)."""
    return np.fft.rfftfreq(n=n_fft, d=1.0 / sr)


def frames_to_samples(frames, *, hop_length=512, n_fft=None):
    offset = 0
    if n_fft is not None:
        offset = int(n_fft // 2)
    return (np.asanyarray(frames) * hop_length + offset).astype(int)


def samples_to_time(samples, *, sr=22050):
    return np.asanyarray(samples) / float(sr)


def frames_to_time(frames, *, sr=22050, hop_length=512, n_fft=None):
    """librosa.frames_to_time (manager.py:168)."""
    samples = frames_to_samples(frames, hop_length=hop_length, n_fft=n_fft)
    return samples_to_time(samples, sr=sr)


def times_like(X, *, sr=22050, hop_length=512, n_fft=None, axis=-1):
    """librosa.times_like (manager.py:194)."""
    if np.isscalar(X):
        frames = np.arange(X)
    else:
        frames = np.arange(X.shape[axis])
    return frames_to_time(frames, sr=sr, hop_length=hop_length, n_fft=n_fft)


def hz_to_mel(frequencies, *, htk=False):
    frequencies = np.asanyarray(frequencies)
    if htk:
        return 2595.0 * np.log10(1.0 + frequencies / 700.0)
    f_min = 0.0
    f_sp = 200.0 / 3
    mels = (frequencies - f_min) / f_sp
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if frequencies.ndim:
        log_t = frequencies >= min_log_hz
        mels[log_t] = min_log_mel + np.log(frequencies[log_t] / min_log_hz) / logstep
    elif frequencies >= min_log_hz:
        mels = min_log_mel + np.log(frequencies / min_log_hz) / logstep
    return mels


def mel_to_hz(mels, *, htk=False):
    mels = np.asanyarray(mels)
    if htk:
        return 700.0 * (10.0 ** (mels / 2595.0) - 1.0)
    f_min = 0.0
    f_sp = 200.0 / 3
    freqs = f_min + f_sp * mels
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if mels.ndim:
        log_t = mels >= min_log_mel
        freqs[log_t] = min_log_hz * np.exp(logstep * (mels[log_t] - min_log_mel))
    elif mels >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (mels - min_log_mel))
    return freqs


def mel_frequencies(n_mels=128, *, fmin=0.0, fmax=11025.0, htk=False):
    min_mel = hz_to_mel(fmin, htk=htk)
    max_mel = hz_to_mel(fmax, htk=htk)
    mels = np.linspace(min_mel, max_mel, n_mels)
    return mel_to_hz(mels, htk=htk)


# ----------------------------------------------------------------------------
# filters
# ----------------------------------------------------------------------------
def _mel(*, sr, n_fft, n_mels=128, fmin=0.0, fmax=None, htk=False, norm="slaney",
         dtype=np.float32):
    """librosa.filters.mel -- Slaney-scale, Slaney-area-normalised triangles,
    stored as float32 (called from melspectrogram; manager.py:219-222)."""
    if fmax is None:
        fmax = float(sr) / 2
    n_mels = int(n_mels)
    weights = np.zeros((n_mels, int(1 + n_fft // 2)), dtype=dtype)
    fftfreqs = fft_frequencies(sr=sr, n_fft=n_fft)
    mel_f = mel_frequencies(n_mels + 2, fmin=fmin, fmax=fmax, htk=htk)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    if isinstance(norm, str):
        if norm == "slaney":
            enorm = 2.0 / (mel_f[2: n_mels + 2] - mel_f[:n_mels])
            weights *= enorm[:, np.newaxis]
        else:
            raise ParameterError(f"Unsupported norm={norm}")
    elif norm is not None:
        raise ParameterError("only norm='slaney' or None are restated")
    return weights


def _get_window(window, Nx, *, fftbins=True):
    """librosa.filters.get_window -> scipy.signal.get_window (periodic)."""
    if callable(window):
        return window(Nx)
    if isinstance(window, (str, tuple)) or np.isscalar(window):
        return scipy.signal.get_window(window, Nx, fftbins=fftbins)
    if isinstance(window, (np.ndarray, list)):
        if len(window) == Nx:
            return np.asarray(window)
        raise ParameterError(f"Window size mismatch: {len(window)} != {Nx}")
    raise ParameterError(f"Invalid window specification: {window!r}")


filters = types.ModuleType("librosa.filters")
filters.mel = _mel
filters.get_window = _get_window


# ----------------------------------------------------------------------------
# core spectrum
# ----------------------------------------------------------------------------
def stft(y, *, n_fft=2048, hop_length=None, win_length=None, window="hann",
         center=True, dtype=None, pad_mode="constant", out=None):
    """librosa.stft (dsp.py:216-224, manager.py:184-187).

    periodic window zero-padded (centred) to n_fft; centre padding n_fft//2 with
    ``pad_mode``; frames at multiples of hop; unnormalised rfft.
    """
    y = np.asarray(y)
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = int(win_length // 4)
    elif hop_length <= 0 or int(hop_length) != hop_length:
        raise ParameterError(f"hop_length={hop_length} must be a positive integer")
    fft_window = _get_window(window, win_length, fftbins=True)
    fft_window = _pad_center(fft_window, size=n_fft)
    if center:
        if pad_mode in ("wrap", "maximum", "mean", "median", "minimum"):
            raise ParameterError(f"pad_mode='{pad_mode}' is not supported by librosa.stft")
        if n_fft > y.shape[-1]:
            import warnings
            warnings.warn(f"n_fft={n_fft} is too large for input signal of length={y.shape[-1]}")
        padding = [(0, 0)] * y.ndim
        padding[-1] = (n_fft // 2, n_fft // 2)
        y = np.pad(y, padding, mode=pad_mode)
    else:
        if n_fft > y.shape[-1]:
            raise ParameterError(
                f"n_fft={n_fft} is too large for uncentered analysis of input signal of length={y.shape[-1]}"
            )
    y_frames = _frame(y, frame_length=n_fft, hop_length=hop_length)  # (..., n_fft, T)
    if dtype is None:
        dtype = np.complex64 if y.dtype == np.float32 else np.complex128
    fw = fft_window.reshape((-1, 1))
    D = np.fft.rfft(fw * y_frames, axis=-2)
    return D.astype(dtype, copy=False)


def _spectrogram(*, y=None, S=None, n_fft=2048, hop_length=512, power=1, win_length=None,
                 window="hann", center=True, pad_mode="constant"):
    if S is not None:
        if n_fft is None or n_fft // 2 + 1 != S.shape[-2]:
            n_fft = 2 * (S.shape[-2] - 1)
    else:
        if n_fft is None:
            raise ParameterError(f"Unable to compute spectrogram with n_fft={n_fft}")
        if y is None:
            raise ParameterError("Input signal must be provided to compute a spectrogram")
        S = np.abs(stft(y, n_fft=n_fft, hop_length=hop_length, win_length=win_length,
                        center=center, window=window, pad_mode=pad_mode)) ** power
    return S, n_fft


def power_to_db(S, *, ref=1.0, amin=1e-10, top_db=80.0):
    """librosa.power_to_db (manager.py:223; inside spectral_contrast)."""
    S = np.asarray(S)
    if amin <= 0:
        raise ParameterError("amin must be strictly positive")
    if np.issubdtype(S.dtype, np.complexfloating):
        magnitude = np.abs(S)
    else:
        magnitude = S
    if callable(ref):
        ref_value = ref(magnitude)
    else:
        ref_value = np.abs(ref)
    log_spec = 10.0 * np.log10(np.maximum(amin, magnitude))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref_value))
    if top_db is not None:
        if top_db < 0:
            raise ParameterError("top_db must be non-negative")
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


def amplitude_to_db(S, *, ref=1.0, amin=1e-5, top_db=80.0):
    S = np.asarray(S)
    magnitude = np.abs(S)
    if callable(ref):
        ref_value = ref(magnitude)
    else:
        ref_value = np.abs(ref)
    power = np.square(magnitude, out=magnitude.copy())
    return power_to_db(power, ref=ref_value ** 2, amin=amin ** 2, top_db=top_db)


def zero_crossings(y, *, threshold=1e-10, ref_magnitude=None, pad=True, zero_pos=True, axis=-1):
    if callable(ref_magnitude):
        threshold = threshold * ref_magnitude(np.abs(y))
    elif ref_magnitude is not None:
        threshold = threshold * ref_magnitude
    yi = np.array(y, copy=True)
    if threshold > 0:
        yi[np.abs(yi) <= threshold] = 0
    if zero_pos:
        sign = np.signbit(yi)
    else:
        sign = np.sign(yi)
    a = np.swapaxes(sign, axis, -1)
    z = a[..., 1:] != a[..., :-1]
    padw = [(0, 0)] * z.ndim
    padw[-1] = (1, 0)
    z = np.pad(z, padw, mode="constant", constant_values=pad)
    return np.swapaxes(z, axis, -1)


# ----------------------------------------------------------------------------
# librosa.feature
# ----------------------------------------------------------------------------
def _melspectrogram(*, y=None, sr=22050, S=None, n_fft=2048, hop_length=512, win_length=None,
                    window="hann", center=True, pad_mode="constant", power=2.0, **kwargs):
    """librosa.feature.melspectrogram (manager.py:219-222).  With ``S`` given it
    is used as is (``power`` is NOT re-applied)."""
    S, n_fft = _spectrogram(y=y, S=S, n_fft=n_fft, hop_length=hop_length, power=power,
                            win_length=win_length, window=window, center=center,
                            pad_mode=pad_mode)
    mel_basis = _mel(sr=sr, n_fft=n_fft, **kwargs)
    return np.einsum("...ft,mf->...mt", S, mel_basis, optimize=True)


def _mfcc(*, y=None, sr=22050, S=None, n_mfcc=20, dct_type=2, norm="ortho", lifter=0, **kwargs):
    """librosa.feature.mfcc (cepstral.py:106-115)."""
    if S is None:
        S = power_to_db(_melspectrogram(y=y, sr=sr, **kwargs))
    M = scipy.fftpack.dct(S, axis=-2, type=dct_type, norm=norm)[..., :n_mfcc, :]
    if lifter > 0:
        LI = np.sin(np.pi * np.arange(1, 1 + n_mfcc, dtype=M.dtype) / lifter)
        LI = _expand_to(LI, ndim=S.ndim, axes=-2)
        M *= 1 + (lifter / 2) * LI
        return M
    elif lifter == 0:
        return M
    raise ParameterError(f"MFCC lifter={lifter} must be a non-negative number")


def _rms(*, y=None, S=None, frame_length=2048, hop_length=512, center=True,
         pad_mode="constant", dtype=np.float32):
    """librosa.feature.rms (audio/features.py:118-126).  NB float32 inside."""
    if y is not None:
        if center:
            padding = [(0, 0) for _ in range(y.ndim)]
            padding[-1] = (int(frame_length // 2), int(frame_length // 2))
            y = np.pad(y, padding, mode=pad_mode)
        x = _frame(y, frame_length=frame_length, hop_length=hop_length)
        power = np.mean(_abs2(x, dtype=dtype), axis=-2, keepdims=True)
    elif S is not None:
        if S.shape[-2] != frame_length // 2 + 1:
            raise ParameterError(
                "Since S.shape[-2] is {}, frame_length is expected to be {} or {}; found {}".format(
                    S.shape[-2], S.shape[-2] * 2 - 2, S.shape[-2] * 2 - 1, frame_length))
        x = _abs2(S, dtype=dtype)
        x[..., 0, :] *= 0.5
        if frame_length % 2 == 0:
            x[..., -1, :] *= 0.5
        power = 2 * np.sum(x, axis=-2, keepdims=True) / frame_length ** 2
    else:
        raise ParameterError("Either `y` or `S` must be input.")
    return np.sqrt(power)


def _zero_crossing_rate(y, *, frame_length=2048, hop_length=512, center=True, **kwargs):
    """librosa.feature.zero_crossing_rate (audio/features.py:26-71) -- edge padding."""
    if center:
        padding = [(0, 0) for _ in range(y.ndim)]
        padding[-1] = (int(frame_length // 2), int(frame_length // 2))
        y = np.pad(y, padding, mode="edge")
    y_framed = _frame(y, frame_length=frame_length, hop_length=hop_length)
    kwargs["axis"] = -2
    kwargs.setdefault("pad", False)
    crossings = zero_crossings(y_framed, **kwargs)
    return np.mean(crossings, axis=-2, keepdims=True)


def _spectral_contrast(*, y=None, sr=22050, S=None, n_fft=2048, hop_length=512, win_length=None,
                       window="hann", center=True, pad_mode="constant", freq=None, fmin=200.0,
                       n_bands=6, quantile=0.02, linear=False):
    """librosa.feature.spectral_contrast (frequency_domain.py:200-207)."""
    S, n_fft = _spectrogram(y=y, S=S, n_fft=n_fft, hop_length=hop_length, win_length=win_length,
                            window=window, center=center, pad_mode=pad_mode)
    if freq is None:
        freq = fft_frequencies(sr=sr, n_fft=n_fft)
    freq = np.atleast_1d(freq)
    if freq.ndim != 1 or len(freq) != S.shape[-2]:
        raise ParameterError(f"freq.shape mismatch: expected ({S.shape[-2]:d},)")
    if n_bands < 1 or not isinstance(n_bands, (int, np.integer)):
        raise ParameterError("n_bands must be a positive integer")
    if not 0.0 < quantile < 1.0:
        raise ParameterError("quantile must lie in the range (0, 1)")
    if fmin <= 0:
        raise ParameterError("fmin must be a positive number")
    octa = np.zeros(n_bands + 2)
    octa[1:] = fmin * (2.0 ** np.arange(0, n_bands + 1))
    if np.any(octa[:-1] >= 0.5 * sr):
        raise ParameterError("Frequency band exceeds Nyquist. Reduce either fmin or n_bands.")
    shape = list(S.shape)
    shape[-2] = n_bands + 1
    valley = np.zeros(shape)
    peak = np.zeros_like(valley)
    for k, (f_low, f_high) in enumerate(zip(octa[:-1], octa[1:])):
        current_band = np.logical_and(freq >= f_low, freq <= f_high)
        idx = np.flatnonzero(current_band)
        if k > 0:
            current_band[idx[0] - 1] = True
        if k == n_bands:
            current_band[idx[-1] + 1:] = True
        sub_band = S[..., current_band, :]
        if k < n_bands:
            sub_band = sub_band[..., :-1, :]
        idx = np.rint(quantile * np.sum(current_band))
        idx = int(np.maximum(idx, 1))
        sortedr = np.sort(sub_band, axis=-2)
        valley[..., k, :] = np.mean(sortedr[..., :idx, :], axis=-2)
        peak[..., k, :] = np.mean(sortedr[..., -idx:, :], axis=-2)
    if linear:
        return peak - valley
    return power_to_db(peak) - power_to_db(valley)


feature = types.ModuleType("librosa.feature")
feature.melspectrogram = _melspectrogram
feature.mfcc = _mfcc
feature.rms = _rms
feature.zero_crossing_rate = _zero_crossing_rate
feature.spectral_contrast = _spectral_contrast


# ----------------------------------------------------------------------------
# signal generators used by the reference's own test fixtures
# ----------------------------------------------------------------------------
def tone(frequency, *, sr=22050, length=None, duration=None, phi=None):
    if length is None:
        length = duration * sr
    if phi is None:
        phi = -np.pi * 0.5
    return np.cos(2 * np.pi * frequency * np.arange(int(length)) / sr + phi)


def chirp(*, fmin, fmax, sr=22050, length=None, duration=None, linear=False, phi=None):
    """librosa.chirp (tests/test_features_manager.py:28)."""
    period = 1.0 / sr
    if length is None:
        duration_ = duration
    else:
        duration_ = period * length
    if phi is None:
        phi = -np.pi * 0.5
    method = "linear" if linear else "logarithmic"
    y = scipy.signal.chirp(np.arange(int(np.ceil(duration_ * sr))) / sr, fmin, duration_, fmax,
                           method=method, phi=phi / np.pi * 180)
    return y


def get_duration(*, y=None, sr=22050, **_):
    return float(y.shape[-1]) / sr


def note_to_hz(note, **_):
    table = {"C": -9, "D": -7, "E": -5, "F": -4, "G": -2, "A": 0, "B": 2}
    import re
    m = re.match(r"^([A-Ga-g])([#b]*)(-?\d+)?$", note)
    if not m:
        raise ParameterError(f"Improper note format: {note}")
    semis = table[m.group(1).upper()] + m.group(2).count("#") - m.group(2).count("b")
    octave = int(m.group(3)) if m.group(3) else 0
    midi = 69 + semis + 12 * (octave - 4)
    return 440.0 * 2.0 ** ((midi - 69) / 12.0)


def _unsupported(name):
    def fn(*a, **k):
        raise NotImplementedError(f"librosa.{name} is outside the hot path and not restated in the shim")
    fn.__name__ = name
    return fn


load = _unsupported("load")
cqt = _unsupported("cqt")
pyin = _unsupported("pyin")
yin = _unsupported("yin")
effects = types.ModuleType("librosa.effects")
for _n in ("hpss", "split", "pitch_shift", "time_stretch", "trim"):
    setattr(effects, _n, _unsupported("effects." + _n))
onset = types.ModuleType("librosa.onset")
onset.onset_detect = _unsupported("onset.onset_detect")
onset.onset_strength = _unsupported("onset.onset_strength")
core = types.ModuleType("librosa.core")
core.stft = stft
