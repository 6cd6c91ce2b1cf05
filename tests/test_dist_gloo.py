"""world_size-2 gloo test of the N>1 path (sygnals_b200/dist.py) on CPU: block partition + halo slicing + the final
all-gather must reproduce the single-rank result exactly.  Compute runs on the CPU *emulator build of the same kernel
sources* (test infrastructure, tests/emu) because this container has no GPU; the sharding/gather code under test is the
product's."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from sygnals_b200 import dist as sdist  # noqa: E402

FEATS = ["mfcc", "spectral_centroid", "rms_energy", "crest_factor"]


def test_shard_range_partitions_exactly():
    for n in (0, 1, 5, 16, 36000):
        for world in (1, 2, 3, 8):
            blocks = [sdist.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_plan_halo_covers_every_segment():
    from backends import get_engine
    lib = get_engine("emu").lib
    total, sr = 10 * 1000 + 321, 1000
    whole = sdist.plan_segments(total, sr, 2.0, 0.5, True, None, 0, 1, lib)
    for world in (2, 3):
        seen = 0
        for r in range(world):
            p = sdist.plan_segments(total, sr, 2.0, 0.5, True, None, r, world, lib)
            assert p.n_units == whole.n_units and p.counts.sum() == whole.n_units
            # relative starts + the rank's base reproduce the global table, and the slice holds every valid sample
            np.testing.assert_array_equal(p.starts + p.sample_begin, whole.starts[p.u0:p.u1])
            np.testing.assert_array_equal(p.valid, whole.valid[p.u0:p.u1])
            assert (p.starts + p.valid <= p.sample_end - p.sample_begin).all()
            seen += p.n_local
        assert seen == whole.n_units


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from backends import get_engine
    from sygnals_b200.utils import synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        eng = get_engine("emu")
        sr = 8000
        y = synth.long_signal(int(4.3 * sr), sr, seed=77)
        r = sdist.segment_features_sharded(y, sr, 1.0, FEATS, overlap_ratio=0.5, frame_length=256, hop_length=128,
                                           feature_params={"mfcc": {"n_mels": 20, "n_mfcc": 5}}, engine=eng)
        q.put((rank, r["names"], r["features"], r["plan"].n_local))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gather_equals_single_rank():
    import torch.multiprocessing as mp
    from backends import get_engine
    from sygnals_b200.utils import synth
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    eng = get_engine("emu")
    sr = 8000
    y = synth.long_signal(int(4.3 * sr), sr, seed=77)
    one = sdist.segment_features_sharded(y, sr, 1.0, FEATS, overlap_ratio=0.5, frame_length=256, hop_length=128,
                                         feature_params={"mfcc": {"n_mels": 20, "n_mfcc": 5}}, rank=0, world=1, engine=eng)
    assert one["features"].shape[0] == 9
    assert sorted(g[3] for g in got) == [4, 5]
    for rank, names, feats, _ in got:
        assert names == one["names"]
        np.testing.assert_array_equal(feats, one["features"])       # same kernels, same units -> bit-identical after gather


def _scaler_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from sygnals_b200.core.ml_utils import scaling
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X = _scaler_matrix()
        a, b = sdist.shard_range(X.shape[0], rank, world)
        res = {}
        for kind, params in (("standard", {}), ("minmax", {"feature_range": (-1.0, 2.0)}), ("robust", {"quantile_range": (10.0, 80.0)})):
            Y, sc = scaling.apply_scaling(torch.from_numpy(X[a:b]), kind, params)
            res[kind] = (Y.numpy(), sc.scale_, {"standard": sc.mean_, "minmax": sc.min_, "robust": sc.center_}[kind])
        q.put((rank, a, b, res))
    finally:
        dist.destroy_process_group()


def _scaler_matrix():
    rng = np.random.default_rng(21)
    X = rng.standard_normal((101, 6)) * np.array([1.0, 10.0, 1e-3, 5.0, 1.0, 2.0]) + np.array([0.0, 3.0, 1.0, -2.0, 7.0, 0.0])
    X[:, 4] = 7.0                                   # constant column -> scale 1 (sklearn _handle_zeros_in_scale)
    X[rng.integers(0, 101, 9), 1] = np.nan          # NaNs are ignored in fit, kept in transform
    return X


def test_sharded_scaler_equals_sklearn_two_ranks():
    import torch.multiprocessing as mp
    from sklearn.preprocessing import MinMaxScaler, RobustScaler, StandardScaler
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_scaler_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=240) for _ in procs], key=lambda g: g[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    X = _scaler_matrix()
    ref = {"standard": StandardScaler().fit(X), "minmax": MinMaxScaler(feature_range=(-1.0, 2.0)).fit(X),
           "robust": RobustScaler(quantile_range=(10.0, 80.0)).fit(X)}
    for kind, sk in ref.items():
        want = sk.transform(X)
        have = np.concatenate([g[3][kind][0] for g in got])
        np.testing.assert_allclose(have, want, rtol=1e-12, atol=1e-12, equal_nan=True)
        for g in got:
            np.testing.assert_allclose(g[3][kind][1], sk.scale_, rtol=1e-12)
            np.testing.assert_allclose(g[3][kind][2], {"standard": getattr(sk, "mean_", None), "minmax": getattr(sk, "min_", None),
                                                       "robust": getattr(sk, "center_", None)}[kind], rtol=1e-12, atol=1e-12)


def _agg_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from backends import get_engine
    from sygnals_b200.utils import synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        eng = get_engine("emu")
        sr = 8000
        y = synth.long_signal(int(5.7 * sr), sr, seed=78)
        r = sdist.segment_features_sharded(y, sr, 1.0, FEATS, overlap_ratio=0.5, frame_length=256, hop_length=128,
                                           feature_params={"mfcc": {"n_mels": 20, "n_mfcc": 5}}, engine=eng,
                                           aggregation={"rms_energy": "max", "mfcc_0": "median"})
        q.put((rank, r["features"]))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_aggregated_vectors_equal_single_rank():
    """The bench's strong-scaling path (run_shard with aggregation -> gather of [segments, rows] float64) on two gloo ranks."""
    import torch.multiprocessing as mp
    from backends import get_engine
    from sygnals_b200.utils import synth
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_agg_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    eng = get_engine("emu")
    sr = 8000
    y = synth.long_signal(int(5.7 * sr), sr, seed=78)
    one = sdist.segment_features_sharded(y, sr, 1.0, FEATS, overlap_ratio=0.5, frame_length=256, hop_length=128,
                                         feature_params={"mfcc": {"n_mels": 20, "n_mfcc": 5}}, rank=0, world=1, engine=eng,
                                         aggregation={"rms_energy": "max", "mfcc_0": "median"})
    assert one["features"].dtype == np.float64 and one["features"].shape == (12, 8)
    for _, feats in got:
        np.testing.assert_array_equal(feats, one["features"])
