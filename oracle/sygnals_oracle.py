"""
TEST INFRASTRUCTURE ONLY -- float64 numpy restatement of the Sygnals
segment->features hot path.  Self-contained (numpy/scipy + ``librosa_shim``) so
it travels to the GPU box, where ``/root/reference`` does not exist.

Every function cites the reference file:line (relative to /root/reference) it
follows; loops are kept per-frame where the reference loops per frame, so that
timing this module is an honest stand-in ("port") for the reference CPU path.
Validated against the unmodified reference by
``tests/test_oracle_vs_reference.py`` and pinned by ``tests/golden/*.npz``.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import scipy.signal

from . import librosa_shim as librosa

_EPSILON = np.finfo(np.float64).eps  # frequency_domain.py:21, time_domain.py:21

# manager.py:38-69 (names only; the oracle restates the in-scope subset + the
# cheap adjacent ones so that ordering of 'all' can be checked)
FRAME_BASED = ("mean_amplitude", "std_dev_amplitude", "skewness", "kurtosis",
               "peak_amplitude", "crest_factor", "signal_entropy",
               "zero_crossing_rate", "rms_energy", "hnr", "jitter", "shimmer")
SPECTRUM_BASED = ("spectral_centroid", "spectral_bandwidth", "spectral_flatness",
                  "spectral_rolloff", "dominant_frequency")
SPECTROGRAM_BASED = ("spectral_contrast",)
MELSPEC_BASED = ("mfcc",)
ALL_KNOWN_FEATURES = set(FRAME_BASED) | set(SPECTRUM_BASED) | set(SPECTROGRAM_BASED) | set(MELSPEC_BASED)


class FeatureExtractionError(Exception):
    """manager.py:72-74"""


# ----------------------------------------------------------------------------
# segmentation.py:25-117
# ----------------------------------------------------------------------------
def segment_table(total_samples: int, sr: int, segment_length_sec: float, overlap_ratio: float = 0.0,
                  pad: bool = True, min_segment_length_sec: Optional[float] = None
                  ) -> Tuple[int, int, List[Tuple[int, int]]]:
    """Integer boundary arithmetic of ``segment_fixed_length`` (segmentation.py:62-114).

    Returns (segment_length_samples, hop_length_samples, [(start, valid_len), ...]).
    """
    if segment_length_sec <= 0:
        raise ValueError("segment_length_sec must be positive.")
    if not 0.0 <= overlap_ratio < 1.0:
        raise ValueError("overlap_ratio must be between 0.0 and < 1.0.")
    seg = int(segment_length_sec * sr)                       # :62
    if seg == 0:
        return 0, 0, []                                       # :63-65
    hop = max(1, int(seg * (1.0 - overlap_ratio)))            # :67-69
    min_samples = int(min_segment_length_sec * sr) if min_segment_length_sec is not None else 0  # :71
    table: List[Tuple[int, int]] = []
    start = 0
    while start < total_samples:                              # :81
        end = start + seg
        orig = min(end, total_samples) - start                # :84
        if min_samples > 0 and orig < min_samples:            # :88
            pass
        elif end > total_samples:                             # :90
            if pad:
                table.append((start, orig))                   # :91-94 zero-padded tail
        else:
            table.append((start, seg))                        # :99-100
        start += hop                                          # :110
        if not pad and start + seg > total_samples:           # :113-114
            break
    return seg, hop, table


def segment_fixed_length(y: np.ndarray, sr: int, segment_length_sec: float, overlap_ratio: float = 0.0,
                         pad: bool = True, min_segment_length_sec: Optional[float] = None) -> List[np.ndarray]:
    """segmentation.py:25-117."""
    if y.ndim != 1:
        raise ValueError("Input signal y must be 1D.")
    seg, _hop, table = segment_table(len(y), sr, segment_length_sec, overlap_ratio, pad, min_segment_length_sec)
    out = []
    for start, valid in table:
        s = y[start:start + valid]
        if valid < seg:
            s = np.pad(s, (0, seg - valid), mode="constant")
        out.append(s.astype(np.float64, copy=False))
    return out


# ----------------------------------------------------------------------------
# dsp.py
# ----------------------------------------------------------------------------
def compute_stft(y, n_fft=2048, hop_length=None, win_length=None, window="hann", center=True,
                 pad_mode="constant"):
    """dsp.py:167-229."""
    if y.ndim != 1:
        raise ValueError("Input data must be a 1D array.")
    D = librosa.stft(y=y, n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window,
                     center=center, pad_mode=pad_mode)
    return D.astype(np.complex128, copy=False)


def compute_psd_welch(x, fs=1.0, window="hann", nperseg=None, noverlap=None, nfft=None,
                      detrend="constant", scaling="density"):
    """dsp.py:495-560 (scipy.signal.welch, one-sided)."""
    if x.ndim != 1:
        raise ValueError("Input data must be a 1D array.")
    f, p = scipy.signal.welch(x, fs=fs, window=window, nperseg=nperseg, noverlap=noverlap, nfft=nfft,
                              detrend=detrend, return_onesided=True, scaling=scaling)
    return f.astype(np.float64, copy=False), p.astype(np.float64, copy=False)


def compute_psd_periodogram(x, fs=1.0, window="hann", nfft=None, detrend="constant", scaling="density"):
    """dsp.py:434-493 (scipy.signal.periodogram, one-sided)."""
    if x.ndim != 1:
        raise ValueError("Input data must be a 1D array.")
    f, p = scipy.signal.periodogram(x, fs=fs, window=window, nfft=nfft, detrend=detrend,
                                    return_onesided=True, scaling=scaling)
    return f.astype(np.float64, copy=False), p.astype(np.float64, copy=False)


def welch_restated(x, fs, nperseg, noverlap=None, scaling="density"):
    """Independent restatement of the scipy Welch recipe used to
    check the kernel's arithmetic step by step: hann periodic, detrend constant,
    |rfft|^2 * scale, one-sided doubling, mean over segments."""
    if noverlap is None:
        noverlap = nperseg // 2
    step = nperseg - noverlap
    nseg = (len(x) - noverlap) // step
    n = np.arange(nperseg)
    w = 0.5 - 0.5 * np.cos(2 * np.pi * n / nperseg)
    scale = 1.0 / (fs * np.sum(w * w)) if scaling == "density" else 1.0 / np.sum(w) ** 2
    acc = np.zeros(nperseg // 2 + 1)
    for s in range(nseg):
        seg = x[s * step: s * step + nperseg]
        seg = seg - seg.mean()
        X = np.fft.rfft(seg * w)
        acc += (X.real ** 2 + X.imag ** 2) * scale
    acc /= nseg
    if nperseg % 2 == 0:
        acc[1:-1] *= 2
    else:
        acc[1:] *= 2
    return np.fft.rfftfreq(nperseg, 1 / fs), acc


# ----------------------------------------------------------------------------
# time_domain.py / audio/features.py
# ----------------------------------------------------------------------------
def peak_amplitude(frame):
    """time_domain.py:128-147"""
    if frame.size == 0:
        return np.float64(0.0)
    return np.max(np.abs(frame))


def crest_factor(frame):
    """time_domain.py:149-184"""
    if frame.size == 0:
        return np.float64(0.0)
    peak = peak_amplitude(frame)
    rms = np.sqrt(np.mean(frame ** 2))
    if rms < _EPSILON:
        return np.float64(0.0)
    return np.float64(peak / rms)


def mean_amplitude(frame):
    """time_domain.py:23-40 (mean of |x|)"""
    if frame.size == 0:
        return np.float64(0.0)
    return np.float64(np.mean(np.abs(frame)))


def std_dev_amplitude(frame):
    """time_domain.py:42-58"""
    if frame.size == 0:
        return np.float64(0.0)
    return np.float64(np.std(frame))


def rms_energy(y=None, *, S=None, frame_length=2048, hop_length=512, center=True, pad_mode="constant", **kw):
    """audio/features.py:73-131 -> librosa.feature.rms (float32 inside) -> float64."""
    if S is None and y is None:
        raise ValueError("Either audio time series 'y' or magnitude spectrogram 'S' must be provided.")
    if y is not None and y.ndim != 1:
        raise ValueError("Input audio data 'y' must be a 1D array.")
    rms = librosa.feature.rms(y=y, S=S, frame_length=frame_length, hop_length=hop_length, center=center,
                              pad_mode=pad_mode, **kw)
    return rms[0].astype(np.float64, copy=False)


def zero_crossing_rate(y, frame_length=2048, hop_length=512, center=True, **kw):
    """audio/features.py:26-71"""
    if y.ndim != 1:
        raise ValueError("Input audio data must be a 1D array.")
    z = librosa.feature.zero_crossing_rate(y, frame_length=frame_length, hop_length=hop_length, center=center, **kw)
    return z[0].astype(np.float64, copy=False)


# ----------------------------------------------------------------------------
# frequency_domain.py
# ----------------------------------------------------------------------------
def spectral_centroid(magnitude_spectrum, frequencies):
    """frequency_domain.py:24-74"""
    if magnitude_spectrum.shape != frequencies.shape:
        raise ValueError("shape mismatch")
    if magnitude_spectrum.size == 0:
        return np.float64(0.0)
    if np.any(magnitude_spectrum < 0):
        magnitude_spectrum = np.abs(magnitude_spectrum)
    s = np.sum(magnitude_spectrum)
    if s < _EPSILON:
        return np.float64(0.0)
    return np.float64(np.sum(frequencies * magnitude_spectrum) / s)


def spectral_bandwidth(magnitude_spectrum, frequencies, centroid=None, p=2):
    """frequency_domain.py:76-145"""
    if p <= 0:
        raise ValueError("Order 'p' for spectral bandwidth must be positive.")
    if magnitude_spectrum.size == 0:
        return np.float64(0.0)
    if np.any(magnitude_spectrum < 0):
        magnitude_spectrum = np.abs(magnitude_spectrum)
    s = np.sum(magnitude_spectrum)
    if s < _EPSILON:
        return np.float64(0.0)
    if centroid is None:
        centroid = spectral_centroid(magnitude_spectrum, frequencies)
    dev = np.abs(frequencies - centroid) ** p
    wsum = np.sum(magnitude_spectrum * dev)
    if wsum < 0:
        wsum = 0.0
    return np.float64((wsum / s) ** (1.0 / p))


def spectral_flatness(magnitude_spectrum):
    """frequency_domain.py:214-271 (geometric/arithmetic mean of the MAGNITUDE spectrum)."""
    if magnitude_spectrum.size == 0:
        return np.float64(0.0)
    if np.any(magnitude_spectrum < 0):
        magnitude_spectrum = np.abs(magnitude_spectrum)
    geometric_mean = np.exp(np.mean(np.log(magnitude_spectrum + _EPSILON)))
    arithmetic_mean = np.mean(magnitude_spectrum)
    if arithmetic_mean < _EPSILON:
        return np.float64(0.0)
    return np.clip(np.float64(geometric_mean / arithmetic_mean), 0.0, 1.0)


def spectral_rolloff(magnitude_spectrum, frequencies, roll_percent=0.85):
    """frequency_domain.py:274-351 (power-based)."""
    if magnitude_spectrum.shape != frequencies.shape:
        raise ValueError("shape mismatch")
    if not 0.0 <= roll_percent <= 1.0:
        raise ValueError("roll_percent must be between 0.0 and 1.0.")
    if magnitude_spectrum.size == 0:
        return np.float64(0.0)
    if np.any(magnitude_spectrum < 0):
        magnitude_spectrum = np.abs(magnitude_spectrum)
    power = magnitude_spectrum ** 2
    total = np.sum(power)
    if total < _EPSILON:
        return frequencies[-1] if frequencies.size > 0 else np.float64(0.0)
    cum = np.cumsum(power)
    idx = np.where(cum >= roll_percent * total)[0]
    if idx.size == 0:
        return frequencies[-1]
    return np.float64(frequencies[idx[0]])


def dominant_frequency(magnitude_spectrum, frequencies):
    """frequency_domain.py:354-386"""
    if magnitude_spectrum.size == 0:
        return np.float64(0.0)
    return np.float64(frequencies[np.argmax(magnitude_spectrum)])


def spectral_contrast(S, sr, n_bands=6, fmin=200.0, freqs=None, **kwargs):
    """frequency_domain.py:147-212"""
    if S.ndim != 2:
        raise ValueError("Input S must be a 2D spectrogram (frequency x time).")
    if np.any(S < 0):
        S = np.abs(S)
    c = librosa.feature.spectral_contrast(S=S, sr=sr, n_bands=n_bands, fmin=fmin, freq=freqs, **kwargs)
    return c.astype(np.float64, copy=False)


# ----------------------------------------------------------------------------
# cepstral.py:20-120
# ----------------------------------------------------------------------------
def mfcc(y=None, sr=None, S=None, n_mfcc=13, dct_type=2, norm="ortho", lifter=0.0, **kwargs):
    if S is None and y is None:
        raise ValueError("Either audio time series 'y' or Mel spectrogram 'S' must be provided.")
    if S is None and sr is None:
        raise ValueError("Sampling rate 'sr' must be provided when calculating MFCCs from time series 'y'.")
    M = librosa.feature.mfcc(y=y, sr=sr, S=S, n_mfcc=n_mfcc, dct_type=dct_type, norm=norm,
                             lifter=float(lifter), **kwargs)
    return M.astype(np.float64, copy=False)


# ----------------------------------------------------------------------------
# manager.py:78-445
# ----------------------------------------------------------------------------
def skewness(frame):
    """time_domain.py:67-97"""
    import scipy.stats
    if frame.size < 2:
        return np.float64(0.0)
    if np.var(frame) < _EPSILON:
        return np.float64(0.0)
    return np.float64(scipy.stats.skew(frame, bias=False))


def kurtosis_val(frame):
    """time_domain.py:99-126"""
    import scipy.stats
    if frame.size < 4:
        return np.float64(0.0)
    if np.var(frame) < _EPSILON:
        return np.float64(0.0)
    return np.float64(scipy.stats.kurtosis(frame, fisher=True, bias=False))


def signal_entropy(frame, num_bins=10):
    """time_domain.py:186-227"""
    import scipy.stats
    if frame.size < 2 or num_bins < 1:
        return np.float64(0.0)
    if np.all(frame == frame[0]):
        return np.float64(0.0)
    counts, _ = np.histogram(frame, bins=num_bins, density=False)
    pk = counts[counts > 0] / frame.size
    return np.float64(scipy.stats.entropy(pk))


_PER_FRAME_TIME = {
    "peak_amplitude": peak_amplitude, "crest_factor": crest_factor,
    "mean_amplitude": mean_amplitude, "std_dev_amplitude": std_dev_amplitude,
    "skewness": skewness, "kurtosis": kurtosis_val, "signal_entropy": signal_entropy,
}
_PER_FRAME_SPEC = {
    "spectral_centroid": spectral_centroid, "spectral_rolloff": spectral_rolloff,
    "spectral_bandwidth": spectral_bandwidth, "spectral_flatness": spectral_flatness,
    "dominant_frequency": dominant_frequency,
}
ORACLE_FEATURES = (set(_PER_FRAME_TIME) | set(_PER_FRAME_SPEC)
                   | {"rms_energy", "zero_crossing_rate", "spectral_contrast", "mfcc"})


def frame_count(n_samples: int, frame_length: int, hop_length: int, center: bool = True) -> int:
    """manager.py:149-157 (predicted) -- NB equals the STFT count only for even frame_length."""
    if center:
        return 1 + n_samples // hop_length
    if n_samples >= frame_length:
        return 1 + (n_samples - frame_length) // hop_length
    return 0


def extract_features(y, sr, features, frame_length=2048, hop_length=512, center=True, window="hann",
                     feature_params: Optional[Dict[str, Dict[str, Any]]] = None,
                     ) -> Dict[str, np.ndarray]:
    """manager.py:78-445, ``output_format='dict_of_arrays'`` branch, for the
    features the engine implements.  Per-frame Python loops are kept (:284, :304-316)."""
    feature_params = feature_params or {}
    if features == ["all"]:
        features = sorted(ALL_KNOWN_FEATURES)
    unknown = [f for f in features if f not in ALL_KNOWN_FEATURES]
    if unknown:
        raise ValueError(f"Unknown feature(s) requested: {unknown}. Available: {sorted(ALL_KNOWN_FEATURES)}")
    if y.ndim != 1:
        raise ValueError("Input audio signal 'y' must be a 1D array.")
    num_frames = frame_count(len(y), frame_length, hop_length, center)
    if num_frames <= 0:
        return {"time": np.array([], dtype=np.float64)}
    results: Dict[str, np.ndarray] = {}
    results["time"] = librosa.frames_to_time(np.arange(num_frames), sr=sr, hop_length=hop_length,
                                             n_fft=frame_length if center else None).astype(np.float64)
    cache: Dict[str, Any] = {}

    def get_mag():
        if "S" not in cache:
            D = librosa.stft(y=y, n_fft=frame_length, hop_length=hop_length, win_length=frame_length,
                             window=window, center=center)                       # :184-187
            if D.shape[1] != len(results["time"]):                               # :189-196
                results["time"] = librosa.times_like(D, sr=sr, hop_length=hop_length,
                                                     n_fft=frame_length).astype(np.float64)
            cache["S"] = np.abs(D).astype(np.float64)                            # :198
            cache["f"] = librosa.fft_frequencies(sr=sr, n_fft=frame_length).astype(np.float64)  # :199
        return cache["S"], cache["f"]

    def get_logmel():
        if "M" not in cache:
            S, _ = get_mag()
            p = feature_params.get("mfcc", {})                                   # :213-217
            S_mel = librosa.feature.melspectrogram(S=S ** p.get("power", 2.0), sr=sr,
                                                   n_mels=p.get("n_mels", 128), fmin=p.get("fmin", 0.0),
                                                   fmax=p.get("fmax", sr / 2.0), n_fft=frame_length)
            cache["M"] = librosa.power_to_db(S_mel, ref=np.max).astype(np.float64)  # :223
        return cache["M"]

    done = set()
    for name in features:
        if name in done:
            continue
        params = feature_params.get(name, {})
        cur = len(results["time"])
        out: List[Tuple[str, np.ndarray]] = []
        try:
            if name == "rms_energy":                                             # :258-263
                out.append((name, rms_energy(y=y, frame_length=frame_length, hop_length=hop_length,
                                             center=center, **params)))
            elif name == "zero_crossing_rate":
                out.append((name, zero_crossing_rate(y=y, frame_length=frame_length, hop_length=hop_length,
                                                     center=center, **params)))
            elif name in _PER_FRAME_TIME:                                        # :265-286
                if center:
                    yp = np.pad(y, frame_length // 2, mode="constant")
                    fr = librosa.util.frame(yp, frame_length=frame_length, hop_length=hop_length)
                else:
                    fr = librosa.util.frame(y, frame_length=frame_length, hop_length=hop_length)
                if fr.shape[1] > cur:
                    fr = fr[:, :cur]
                fn = _PER_FRAME_TIME[name]
                import inspect
                kw = {k: v for k, v in params.items() if k in inspect.signature(fn).parameters}   # :283 user params only
                out.append((name, np.array([fn(fr[:, i], **kw) for i in range(fr.shape[1])], dtype=np.float64)))
            elif name in _PER_FRAME_SPEC:                                        # :289-319
                S, f = get_mag()
                cur = S.shape[1]
                fn = _PER_FRAME_SPEC[name]
                if name == "spectral_bandwidth" and "spectral_centroid" not in results:   # :296-301
                    results["spectral_centroid"] = np.array(
                        [spectral_centroid(S[:, i], f) for i in range(cur)], dtype=np.float64)
                    done.add("spectral_centroid")
                vals = []
                for i in range(cur):
                    if name == "spectral_flatness":
                        vals.append(fn(S[:, i]))
                    elif name == "spectral_bandwidth":
                        vals.append(fn(S[:, i], f, centroid=results["spectral_centroid"][i], **params))
                    else:
                        vals.append(fn(S[:, i], f, **params))
                out.append((name, np.array(vals, dtype=np.float64)))
            elif name == "spectral_contrast":                                    # :322-345
                S, f = get_mag()
                cur = S.shape[1]
                c = spectral_contrast(S=S, sr=sr, freqs=f, **params)
                nb = c.shape[0] - 1
                for i in range(nb):
                    out.append((f"contrast_band_{i}", c[i, :cur]))
                out.append(("contrast_delta", c[nb, :cur]))
            elif name == "mfcc":                                                 # :348-371
                M = get_logmel()
                cur = M.shape[1]
                kw = {k: v for k, v in params.items() if k in ("n_mfcc", "dct_type", "norm", "lifter")}
                c = mfcc(S=M, sr=sr, **kw)
                for i in range(c.shape[0]):
                    out.append((f"mfcc_{i}", c[i, :cur]))
            else:
                raise NotImplementedError(f"oracle does not restate feature '{name}' (out of hot-path scope)")
            for n, a in out:                                                     # :374-389
                if len(a) != cur:
                    if len(a) > cur:
                        a = a[:cur]
                    else:
                        b = np.full(cur, np.nan, dtype=np.float64)
                        b[:len(a)] = a
                        a = b
                results[n] = a
                done.add(n)
            done.add(name)
        except NotImplementedError:
            raise
        except Exception:                                                        # :394-397 feature dropped
            continue
    final_n = len(results["time"])
    final = {"time": results["time"]}
    for n, a in results.items():
        if n != "time" and a.ndim == 1 and len(a) == final_n:
            final[n] = a
    return final


# ----------------------------------------------------------------------------
# ml_utils/formatters.py:51-163 (segment aggregation; "next" row f1)
# ----------------------------------------------------------------------------
def aggregate_frames(frames_matrix: np.ndarray, method: str) -> np.ndarray:
    """formatters.py:28-47: NaN-aware aggregate over axis 0 of a (T, n_features) block."""
    fn = {"mean": np.nanmean, "std": np.nanstd, "median": np.nanmedian, "min": np.nanmin, "max": np.nanmax}[method]
    return fn(frames_matrix, axis=0)


# ----------------------------------------------------------------------------
# ml_utils/formatters.py:28-47, 51-163
# ----------------------------------------------------------------------------
def _nanagg(func):
    """formatters.py:28-37"""
    def wrapper(a):
        if a.size == 0 or np.all(np.isnan(a)):
            return np.nan
        valid = a[~np.isnan(a)]
        if valid.size == 0:
            return np.nan
        return func(valid)
    return wrapper


AGGREGATION_FUNCS = {"mean": _nanagg(np.mean), "std": _nanagg(np.std), "median": _nanagg(np.median),
                     "min": _nanagg(np.min), "max": _nanagg(np.max)}          # formatters.py:39-45


def format_feature_vectors_per_segment(features_dict, segment_indices, aggregation="mean"):
    """formatters.py:51-163, ``output_format='numpy'``: [n_segments, n_features] float64, NaN rows for invalid segments."""
    names = list(features_dict.keys())
    num_frames = len(features_dict[names[0]])
    if isinstance(aggregation, str):
        funcs = {n: AGGREGATION_FUNCS[aggregation] for n in names}
    else:
        funcs = {n: AGGREGATION_FUNCS[aggregation.get(n, "mean")] for n in names}
    out = np.full((len(segment_indices), len(names)), np.nan, dtype=np.float64)
    for i, (s, e) in enumerate(segment_indices):
        if not (0 <= s < num_frames and s < e and e <= num_frames):           # formatters.py:138-147
            continue
        for j, n in enumerate(names):
            out[i, j] = funcs[n](np.asarray(features_dict[n], dtype=np.float64)[s:e])
    return out


# ------------------------------------------------------------------------------------------------ audio ingest
# sygnals/core/audio/io.py:38-102 load_audio -> librosa.load(path, sr=None, mono=True) -> soundfile.read(dtype='float32',
# always_2d=False).T -> librosa.to_mono (np.mean(y, axis=0)) -> .astype(float64) (io.py:94-95).  soundfile / libsndfile are
# third-party (python-soundfile >= 0.12 over libsndfile 1.x; not vendored under /root/reference, not installable offline): the
# integer -> float normalisation below restates libsndfile's published behaviour for float reads of PCM files
# (src/pcm.c: u8 (x - 128) / 128, s16 x / 0x8000, s24 and s32 as 32-bit words / 0x80000000); the mix-down is numpy itself.
PCM_U8, PCM_S16, PCM_S24, PCM_S32, PCM_F32 = 0, 1, 2, 3, 4


def pcm_payload_to_float32(raw, fmt: int, channels: int) -> np.ndarray:
    """Bytes of a WAV data chunk -> float32 [frames, channels] as soundfile.read(dtype='float32', always_2d=True) returns them."""
    raw = np.frombuffer(bytes(raw), dtype=np.uint8) if not isinstance(raw, np.ndarray) else raw.view(np.uint8).reshape(-1)
    bps = {PCM_U8: 1, PCM_S16: 2, PCM_S24: 3, PCM_S32: 4, PCM_F32: 4}[fmt]
    frames = raw.size // (bps * channels)
    raw = raw[: frames * bps * channels]
    if fmt == PCM_U8:
        x = (raw.astype(np.int32) - 128).astype(np.float32) * np.float32(1.0 / 128.0)
    elif fmt == PCM_S16:
        x = raw.view("<i2").astype(np.float32) * np.float32(1.0 / 32768.0)
    elif fmt == PCM_S24:
        b = raw.reshape(-1, 3).astype(np.int32)
        w = (b[:, 0] << 8) | (b[:, 1] << 16) | (b[:, 2] << 24)                  # libsndfile: the sample in the top 24 bits of an int32
        x = w.astype(np.int32).astype(np.float32) * np.float32(1.0 / 2147483648.0)
    elif fmt == PCM_S32:
        x = raw.view("<i4").astype(np.float32) * np.float32(1.0 / 2147483648.0)
    elif fmt == PCM_F32:
        x = raw.view("<f4").copy()
    else:
        raise ValueError(fmt)
    return x.reshape(frames, channels)


def load_audio_payload(raw, fmt: int, channels: int, mono: bool = True) -> np.ndarray:
    """load_audio(...)[0] for a WAV payload: float64, (n,) if mono else (channels, n)."""
    y = pcm_payload_to_float32(raw, fmt, channels).T                              # librosa.load: y = sf_desc.read(...).T
    if mono and y.shape[0] > 1:
        y = np.mean(y, axis=0)                                                    # librosa.to_mono
    elif y.shape[0] == 1:
        y = y[0]
    return y.astype(np.float64)
