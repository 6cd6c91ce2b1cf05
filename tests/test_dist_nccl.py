"""The product's multi-GPU path on REAL GPUs (needs >= 2 devices: `gpurun --gpus 2 -- python -m pytest tests/test_dist_nccl.py -m gpu`):
two NCCL ranks run `dist.segment_features_sharded` (block partition + halo, fused kernels, on-device aggregation, one all-gather)
and must reproduce the single-rank result bit for bit; plus one process driving two engines (per-device kernel
attributes) and the sharded scaler on CUDA tensors over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

FEATS = ["mfcc", "spectral_contrast", "spectral_centroid", "spectral_rolloff", "rms_energy", "crest_factor"]


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_n_gpus() < 2, reason="needs two CUDA devices")


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    from sygnals_b200 import _ffi, dist as sdist
    from sygnals_b200.core.ml_utils import scaling
    from sygnals_b200.utils import synth
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        sr = 44100
        total = int(61.3 * sr)
        eng = _ffi.engine(rank)
        plan = sdist.plan_segments(total, sr, 2.0, 0.5, True, None, rank, world, eng.lib)
        y = torch.empty(plan.sample_end - plan.sample_begin, dtype=torch.float32, device=dev)
        synth.torch_recording_(y, plan.sample_begin, sr, seed=5)              # only this rank's slice of the recording
        vec = sdist.gather_features(sdist.run_shard(y, plan, sr, FEATS, engine=eng, aggregation="mean"), plan)
        frames = sdist.gather_features(sdist.run_shard(y, plan, sr, FEATS, engine=eng), plan)
        # scaler fit over the sharded vectors (one all-reduce of the moments), transform of the local block
        local = sdist.run_shard(y, plan, sr, FEATS, engine=eng, aggregation="mean")
        Y, sc = scaling.apply_scaling(local, "standard", {})
        torch.cuda.synchronize()
        q.put((rank, vec.cpu().numpy(), frames.cpu().numpy(), Y.cpu().numpy(), np.asarray(sc.mean_), np.asarray(sc.scale_)))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@needs2
def test_two_rank_nccl_equals_single_rank():
    import torch
    import torch.multiprocessing as mp
    from sklearn.preprocessing import StandardScaler
    from sygnals_b200 import _ffi, dist as sdist
    from sygnals_b200.utils import synth
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=600) for _ in procs], key=lambda g: g[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    sr = 44100
    total = int(61.3 * sr)
    dev = torch.device("cuda", 0)
    y = torch.empty(total, dtype=torch.float32, device=dev)
    synth.torch_recording_(y, 0, sr, seed=5)
    one_v = sdist.segment_features_sharded(y, sr, 2.0, FEATS, overlap_ratio=0.5, rank=0, world=1, aggregation="mean")["features"].cpu().numpy()
    one_f = sdist.segment_features_sharded(y, sr, 2.0, FEATS, overlap_ratio=0.5, rank=0, world=1)["features"].cpu().numpy()
    assert one_v.shape == (62, 24) and one_f.shape == (62, 24, 173)
    for rank, vec, frames, Y, mean, scale in got:
        np.testing.assert_array_equal(frames, one_f)                          # halo slicing + gather: bit-identical frame features
        np.testing.assert_array_equal(vec, one_v)
    sk = StandardScaler().fit(one_v)
    np.testing.assert_allclose(got[0][4], sk.mean_, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(got[0][5], sk.scale_, rtol=1e-10)
    np.testing.assert_allclose(np.concatenate([g[3] for g in got]), sk.transform(one_v), rtol=1e-9, atol=1e-9)


@pytest.mark.gpu
@needs2
def test_one_process_two_engines():
    """engine(0) and engine(1) in ONE process: every kernel needs its shared-memory opt-in on both devices, and a call must not
    change the caller's current device."""
    import torch
    from sygnals_b200 import batch
    from sygnals_b200.utils import synth
    sr = 44100
    y = synth.long_signal(int(4.5 * sr), sr, seed=3)
    torch.cuda.set_device(0)
    outs = []
    for d in (0, 1, 0, 1):
        yd = torch.from_numpy(y).to(f"cuda:{d}")
        r = batch.segment_features(yd, sr, 2.0, FEATS, overlap_ratio=0.5)
        st = batch.stft_batch(yd[: 3 * 16000].reshape(3, 16000), n_fft=4096, output="complex")
        outs.append((r["features"].cpu().numpy(), st.cpu().numpy()))
        assert torch.cuda.current_device() == 0
    for f, s in outs[1:]:
        np.testing.assert_array_equal(f, outs[0][0])
        np.testing.assert_array_equal(s, outs[0][1])
