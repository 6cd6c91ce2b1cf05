import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sygnals_b200 import batch
from sygnals_b200.utils import synth
n_fft = int(sys.argv[1]) if len(sys.argv) > 1 else 512
clips = torch.empty((4096, 16000), dtype=torch.float32, device="cuda")
synth.torch_mixture_(clips, 16000, seed=2)
for _ in range(3):
    out = batch.stft_batch(clips, n_fft=n_fft, output="magnitude")
torch.cuda.synchronize()
print(out.shape)
