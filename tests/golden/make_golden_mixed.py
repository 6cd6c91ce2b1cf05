"""
Generates tests/golden/mixed_lengths.npz: the UNMODIFIED reference (/root/reference, librosa = oracle/librosa_shim.py) on transform
lengths that are NOT powers of two -- the 25 ms / 10 ms speech framing (frame_length 400), frame_length 1000 (the value the
reference's own tests use), an odd length, and the 25 600-sample second of BASELINE config 5 (periodogram + Welch-1000).
Run in the build container only:

    python tests/golden/make_golden_mixed.py
"""
import logging
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from sygnals_b200.utils import synth  # noqa: E402

warnings.simplefilter("ignore")
logging.disable(logging.CRITICAL)
ref_loader.load_reference()
from sygnals.core.dsp import compute_psd_periodogram, compute_psd_welch, compute_stft  # noqa: E402
from sygnals.core.features.manager import extract_features  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
ENV_FEATURES = ["mfcc", "spectral_contrast", "spectral_centroid", "spectral_rolloff", "rms_energy", "crest_factor"]


def checksum(x):
    x = np.asarray(x, dtype=np.float64).ravel()
    return np.array([x.sum(), np.abs(x).sum(), (x * np.arange(1, x.size + 1)).sum()])


def stack(d, skip=("time",)):
    names = [k for k in d if k not in skip]
    return names, np.stack([d[k] for k in names])


def main():
    d = {}
    sr = 16000
    clip = synth.mixture(8000, sr, seed=202)
    d["stft_in_checksum"] = checksum(clip)
    d["D_400_160"] = compute_stft(clip.astype(np.float64), n_fft=400, hop_length=160).astype(np.complex64)
    d["D_1000_250_reflect"] = compute_stft(clip.astype(np.float64), n_fft=1000, hop_length=250, pad_mode="reflect").astype(np.complex64)
    d["D_441_147_nocenter"] = compute_stft(clip.astype(np.float64), n_fft=441, hop_length=147, center=False).astype(np.complex64)
    d["D_1200_win900_300"] = compute_stft(clip.astype(np.float64), n_fft=1200, hop_length=300, win_length=900).astype(np.complex64)

    clips = synth.clip_batch(8, 16000, 16000, seed=606, edges=True)
    rows = []
    for c in clips:
        r = extract_features(c.astype(np.float64), 16000, ["mfcc", "rms_energy"], frame_length=400, hop_length=160,
                             feature_params={"mfcc": {"n_mels": 40}}, output_format="dict_of_arrays")
        names, m = stack(r)
        rows.append(m)
    d["speech400_names"] = np.array(names)
    d["speech400_rows"] = np.stack(rows)
    d["speech400_in_checksum"] = checksum(clips)

    sr = 44100
    y = synth.long_signal(sr, sr, seed=707)
    r = extract_features(y.astype(np.float64), sr, ENV_FEATURES, frame_length=1000, hop_length=250, output_format="dict_of_arrays")
    names, m = stack(r)
    d["env1000_names"] = np.array(names)
    d["env1000_rows"] = m
    d["env1000_in_checksum"] = checksum(y)

    sr = 25600
    x = synth.long_signal(sr, sr, seed=808, block_sec=0.5)
    d["psd_in_checksum"] = checksum(x)
    f, p = compute_psd_periodogram(x.astype(np.float64), fs=sr, window="hann")
    d["periodogram_25600"] = p
    f, p = compute_psd_welch(x.astype(np.float64), fs=sr, window="hann", nperseg=1000, noverlap=500)
    d["welch_1000_500"] = p
    f, p = compute_psd_welch(x.astype(np.float64), fs=sr, window="hamming", nperseg=945, noverlap=100, nfft=1890, scaling="spectrum")
    d["welch_945_100_1890_spectrum"] = p
    np.savez_compressed(os.path.join(OUT, "mixed_lengths.npz"), **d)
    print("mixed_lengths.npz", os.path.getsize(os.path.join(OUT, "mixed_lengths.npz")))


if __name__ == "__main__":
    main()
