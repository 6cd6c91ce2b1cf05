"""Device timings of the mixed-radix kernels (lengths that are not powers of two) next to their power-of-two neighbours:
STFT 400/160 vs 512/160 on 4096 x 1 s clips, speech MFCC 400/160 vs 512/160 on 20 000 clips, periodogram of 25 600-sample
seconds vs Welch-1024 on 7 680 channel-seconds.  CUDA events, 3 warm-ups, 5 timed repetitions, JSON lines."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sygnals_b200 import batch
from sygnals_b200.utils import synth


def timed(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


clips = torch.empty((4096, 16000), dtype=torch.float32, device="cuda")
synth.torch_mixture_(clips, 16000, seed=2)
for n_fft in (400, 512, 1000, 1024):
    ms = timed(lambda: batch.stft_batch(clips, n_fft=n_fft, hop_length=160, output="magnitude"))
    print(json.dumps({"what": "stft", "n_fft": n_fft, "hop": 160, "clips": 4096, "ms": round(ms, 4)}))
sp = torch.empty((20000, 16000), dtype=torch.float32, device="cuda")
synth.torch_mixture_(sp, 16000, seed=3)
for fl in (400, 512):
    ms = timed(lambda: batch.extract_features_batch(sp, 16000, ["mfcc"], frame_length=fl, hop_length=160, feature_params={"mfcc": {"n_mels": 40}}))
    print(json.dumps({"what": "mfcc13/40", "frame_length": fl, "hop": 160, "clips": 20000, "ms": round(ms, 4), "audio_s_per_s": round(20000 / ms * 1e3)}))
fs = 25600
y = torch.empty((7680, fs), dtype=torch.float32, device="cuda")
synth.torch_mixture_(y, fs, seed=5)
ms = timed(lambda: batch.psd_welch_batch(y, fs, nperseg=1024, noverlap=512))
print(json.dumps({"what": "welch", "nperseg": 1024, "units": 7680, "ms": round(ms, 4)}))
ms = timed(lambda: batch.psd_welch_batch(y, fs, nperseg=1000, noverlap=500))
print(json.dumps({"what": "welch", "nperseg": 1000, "units": 7680, "ms": round(ms, 4)}))
ms = timed(lambda: batch.psd_welch_batch(y, fs, nperseg=fs, noverlap=0))
print(json.dumps({"what": "periodogram", "nfft": fs, "units": 7680, "ms": round(ms, 4), "channel_s_per_s": round(7680 / ms * 1e3)}))
