import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sygnals_b200 import batch
from sygnals_b200.utils import synth
clips = torch.empty((30000, 16000), dtype=torch.float32, device="cuda")
synth.torch_mixture_(clips, 16000, seed=3)
for _ in range(3):
    names, out = batch.extract_features_batch(clips, 16000, ["mfcc"], 512, 160, feature_params={"mfcc": {"n_mels": 40}})
torch.cuda.synchronize()
print(out.shape)
