"""Inputs of the golden cases (regenerated from seeds; must match tests/golden/make_golden.py)."""
import os

import numpy as np

from sygnals_b200.utils import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CFG4_FEATURES = ["mfcc", "spectral_contrast", "spectral_centroid", "spectral_rolloff", "rms_energy", "crest_factor"]


def checksum(x):
    x = np.asarray(x, dtype=np.float64).ravel()
    return np.array([x.sum(), np.abs(x).sum(), (x * np.arange(1, x.size + 1)).sum()])


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def cfg1_input():
    return synth.mixture(10 * 22050, 22050, seed=101), 22050


def cfg2_input():
    sr = 16000
    return np.stack([synth.mixture(8000, sr, seed=202), synth.edge_clip("impulse", 8000, sr)]), sr


def cfg3_input():
    return synth.clip_batch(10, 16000, 16000, seed=303, edges=True), 16000


def cfg4_input():
    sr = 44100
    return synth.long_signal(int(7.3 * sr), sr, seed=404), sr


def cfg5_input():
    sr = 25600
    return np.stack([synth.long_signal(2 * sr, sr, seed=505 + c, block_sec=0.5) for c in range(3)]), sr


def stack_rows(d, names):
    return np.stack([d[str(n)] for n in names])
