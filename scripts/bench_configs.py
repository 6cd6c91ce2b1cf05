#!/usr/bin/env python
"""Secondary measurements for BASELINE.json configs 2, 3 and 5 (the default bench.py line is config 4): device-resident inputs,
CUDA events, >= 3 warm-ups, median of the timed runs; algorithmic bytes per SURVEY.md 8(d) (4 B per unique input sample + 4 B
per output float) against the measured HBM copy peak.  One JSON line per point."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sygnals_b200 import batch  # noqa: E402
from sygnals_b200.utils import synth  # noqa: E402


def timed(fn, reps=7, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    which = sys.argv[1:] or ["cfg2", "cfg3", "cfg5"]      # add "yardstick" for the torch/torchaudio library lines
    if "cfg2" in which:
        sr, n, L = 16000, 4096, 16000
        clips = torch.empty((n, L), dtype=torch.float32, device="cuda")
        synth.torch_mixture_(clips, sr, seed=2)
        for n_fft in (256, 512, 1024, 2048, 4096, 8192):
            out = batch.stft_batch(clips, n_fft=n_fft, output="magnitude")
            ms = timed(lambda: batch.stft_batch(clips, n_fft=n_fft, output="magnitude"))
            by = 4.0 * clips.numel() + 4.0 * out.numel()
            print(json.dumps({"config": "cfg2 stft magnitude", "n_fft": n_fft, "hop": n_fft // 4, "clips": n, "ms": ms,
                              "audio_s_per_s": n * 1.0 / (ms * 1e-3), "alg_GB": by / 1e9, "GBps": by / ms / 1e6,
                              "hbm_frac": by / ms / 1e6 / peak}), flush=True)
            del out
            if "yardstick" in which:
                # in-image GPU LIBRARY yardstick: torch.stft (cuFFT) + abs on the same clips -- not the reference, not this engine
                win = torch.hann_window(n_fft, periodic=True, device="cuda")
                lib = lambda: torch.stft(clips, n_fft, hop_length=n_fft // 4, win_length=n_fft, window=win, center=True,
                                         pad_mode="constant", return_complex=True).abs()
                ms_l = timed(lib)
                print(json.dumps({"config": "cfg2 yardstick torch.stft(cuFFT).abs()", "n_fft": n_fft, "ms": ms_l,
                                  "audio_s_per_s": n * 1.0 / (ms_l * 1e-3), "engine_speedup": ms_l / ms}), flush=True)
        del clips
    if "cfg3" in which:
        sr, n, L = 16000, 100000, 16000
        clips = torch.empty((n, L), dtype=torch.float32, device="cuda")
        synth.torch_mixture_(clips, sr, seed=3)
        fp = {"mfcc": {"n_mels": 40}}
        _, out = batch.extract_features_batch(clips, sr, ["mfcc"], 512, 160, feature_params=fp)
        ms = timed(lambda: batch.extract_features_batch(clips, sr, ["mfcc"], 512, 160, feature_params=fp))
        by = 4.0 * clips.numel() + 4.0 * out.numel()
        print(json.dumps({"config": "cfg3 speech-commands mfcc13 (n_fft 512 hop 160, 40 mels)", "clips": n, "ms": ms,
                          "audio_s_per_s": n * 1.0 / (ms * 1e-3), "alg_GB": by / 1e9, "GBps": by / ms / 1e6,
                          "hbm_frac": by / ms / 1e6 / peak}), flush=True)
        if "yardstick" in which:
            # library yardstick: torchaudio MFCC (cuFFT STFT -> dense mel matmul -> global-ref dB -> DCT matmul); its dB reference
            # and top_db handling differ slightly from librosa's per-call ref=np.max, so this is a speed yardstick only
            import torchaudio
            tf = torchaudio.transforms.MFCC(sample_rate=sr, n_mfcc=13, log_mels=False,
                                            melkwargs={"n_fft": 512, "hop_length": 160, "n_mels": 40, "center": True, "pad_mode": "constant",
                                                       "norm": "slaney", "mel_scale": "slaney"}).to("cuda")
            def lib():
                for i in range(0, n, 25000):             # 4 chunks: the library materialises complex spectra (257 x 101 x 8 B per clip)
                    tf(clips[i:i + 25000])
            ms_l = timed(lib)
            print(json.dumps({"config": "cfg3 yardstick torchaudio.transforms.MFCC (4 chunks of 25k clips)", "ms": ms_l,
                              "audio_s_per_s": n * 1.0 / (ms_l * 1e-3), "engine_speedup": ms_l / ms}), flush=True)
        del clips, out
    if "cfg5" in which:
        fs, ch, seconds = 25600, 64, 600
        y = torch.empty((ch * seconds, fs), dtype=torch.float32, device="cuda")
        synth.torch_mixture_(y, fs, seed=5)
        psd, st = batch.psd_welch_batch(y, fs, nperseg=1024, noverlap=512)
        ms = timed(lambda: batch.psd_welch_batch(y, fs, nperseg=1024, noverlap=512))
        by = 4.0 * y.numel() + 4.0 * (psd.numel() + 2 * st.shape[0])
        print(json.dumps({"config": "cfg5 machinery welch(1024/512)+rms+crest per channel-second", "units": ch * seconds, "ms": ms,
                          "channel_s_per_s": ch * seconds / (ms * 1e-3), "alg_GB": by / 1e9, "GBps": by / ms / 1e6,
                          "hbm_frac": by / ms / 1e6 / peak}), flush=True)


if __name__ == "__main__":
    main()
