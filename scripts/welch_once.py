import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sygnals_b200 import batch
from sygnals_b200.utils import synth
fs, ch, seconds = 25600, 64, 300
y = torch.empty((ch * seconds, fs), dtype=torch.float32, device="cuda")
synth.torch_mixture_(y, fs, seed=5)
for _ in range(3):
    psd, st = batch.psd_welch_batch(y, fs, nperseg=1024, noverlap=512)
torch.cuda.synchronize()
print(psd.shape, st.shape)
