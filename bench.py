#!/usr/bin/env python
"""bench.py -- throughput of the Sygnals segment->features hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4|cfg3|cfg2|cfg5] [--impl native|reference]

Default workload = BASELINE.json configs[3] ("cfg4", the config the metric "audio-sec/sec (MFCC+spectral feats)" names):
10 h of synthetic 44.1 kHz audio per GPU, fixed-length 2 s segments with 50 % overlap (36 000 segments),
MFCC(13) + spectral contrast(7) + centroid + rolloff + RMS + crest per 2048/512 frame -> float32 [36000, 24, 173].
One "step" = one pass of the whole path over the resident 10 h buffer.  Weak scaling: every rank owns its own 10 h
recording (units shard with no data-path collective); value = all ranks' audio-seconds / max-over-ranks device time.

`--impl reference` times the reference's CPU algorithm (oracle/: the reference's own call graph restated on numpy/scipy,
float64, incl. its per-frame Python loops; librosa itself is not installable offline) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG4_FEATURES = ["mfcc", "spectral_contrast", "spectral_centroid", "spectral_rolloff", "rms_energy", "crest_factor"]

WORKLOADS = {
    # name: dict(sr, unit seconds, hop seconds between units, features, frame_length, hop_length, feature_params)
    "cfg4": dict(desc="env-sound: 10 h @ 44.1 kHz, 2 s segments, 50% overlap, MFCC13+contrast7+centroid+rolloff+rms+crest, n_fft 2048 hop 512",
                 sr=44100, seg_sec=2.0, overlap=0.5, hours=10.0, features=CFG4_FEATURES, fl=2048, hop=512, fp=None),
    "cfg3": dict(desc="speech-commands: 100k x 1 s @ 16 kHz clips, MFCC13, n_fft 512 hop 160, 40 mels",
                 sr=16000, seg_sec=1.0, overlap=0.0, hours=100000 / 3600.0, features=["mfcc"], fl=512, hop=160,
                 fp={"mfcc": {"n_mels": 40}}),
}


_REAL_STDOUT = None


def guard_stdout():
    """stdout carries exactly ONE JSON line.  Native libraries write banners straight to file descriptor 1 (NCCL prints its version
    there whenever NCCL_DEBUG is set), so fd 1 is pointed at stderr for the whole run and the JSON goes out through a saved copy."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_json(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ------------------------------------------------------------------------------------------------ CPU reference arm
def _cpu_worker(args):
    """One oracle call on one segment (runs in a pool worker)."""
    seg, sr, features, fl, hop, fp = args
    from oracle import sygnals_oracle as orc
    import numpy as np
    r = orc.extract_features(seg.astype(np.float64), sr, list(features), frame_length=fl, hop_length=hop, feature_params=fp)
    return sum(float(v.sum()) for k, v in r.items() if k != "time" and v.size)


def cpu_reference_run(workload: str, n_units: int, steps: int, warmup: int, procs: int | None = None):
    """Times the oracle (CPU restatement of the reference path) on `n_units` segments per step with a process pool over
    all usable host cores.  Returns (units_per_sec list per step, cores)."""
    import multiprocessing as mp
    import numpy as np
    from oracle import sygnals_oracle as orc
    from sygnals_b200.utils import synth
    w = WORKLOADS[workload]
    sr = w["sr"]
    cores = procs or len(os.sched_getaffinity(0))
    seg_len = int(w["seg_sec"] * sr)
    seg_hop = max(1, int(seg_len * (1.0 - w["overlap"])))
    total = seg_hop * (n_units - 1) + seg_len
    y = synth.long_signal(total, sr, seed=4321)
    segs = orc.segment_fixed_length(y.astype(np.float64), sr, w["seg_sec"], overlap_ratio=w["overlap"], pad=False)[:n_units]
    assert len(segs) == n_units, (len(segs), n_units)
    jobs = [(s.astype(np.float32), sr, w["features"], w["fl"], w["hop"], w["fp"]) for s in segs]
    times = []
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, jobs[:cores], chunksize=1)                # import + plan warm-up in every worker
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            pool.map(_cpu_worker, jobs, chunksize=max(1, len(jobs) // (cores * 4)))
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    return times, cores


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    # one process per core, one thread per process: numpy/scipy/BLAS thread pools inside 16 forked workers oversubscribe the
    # host and slow the reference down ~8x (measured); the baseline should be as fast as the reference can go
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        os.environ[v] = "1"
    w = WORKLOADS[args.workload]
    cores = len(os.sched_getaffinity(0))
    n_units = args.cpu_units or max(cores * 24, 96)
    times, cores = cpu_reference_run(args.workload, n_units, args.steps, args.warmup)
    unit_audio = w["seg_sec"] * (1.0 - w["overlap"])                    # unique audio seconds per unit
    total_t = sum(times)
    value = n_units * unit_audio * len(times) / total_t
    line = {
        "impl": "reference", "metric": "audio-sec/sec (MFCC+spectral feats)", "value": value, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w['desc']}", "sample": f"{n_units} segments per step"},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port",
                         "sample": f"{n_units} x {w['seg_sec']} s segments per step through oracle.extract_features "
                                   f"(reference call graph on numpy/scipy, float64), fork pool over {cores} cores"},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json(line)
    return 0


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons, pw = [], [], set(), []
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # median over the upper half of the samples (the sampler also sees the idle edges of the region)
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:]
        return {"sm_mhz": load[len(load) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(pw)}


# ------------------------------------------------------------------------------------------------ native arm
def run_native(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from sygnals_b200 import _ffi
    from sygnals_b200.utils import synth

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # host placement for the end-to-end arm: this rank's pinned buffers live on its GPU's NUMA node (no effect on `value`)
    from sygnals_b200.utils import numa
    full_affinity = os.sched_getaffinity(0)
    placement = None if os.environ.get("SYGB200_NO_NUMA_BIND") else numa.bind_to_device_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = WORKLOADS[args.workload]
    sr, fl, hop = w["sr"], w["fl"], w["hop"]
    eng = _ffi.engine(local)
    if args.ws_mb:
        eng.set_workspace_limit(args.ws_mb << 20)
    lib = eng.lib
    hours = args.hours if args.hours else w["hours"]
    total = int(round(hours * 3600 * sr))
    seg_len, seg_hop, starts, valid = lib.segment_table(total, sr, w["seg_sec"], w["overlap"], True, None)
    n_units = len(starts)
    p = _ffi.make_params(lib, sr, w["features"], fl, hop, feature_params=w["fp"])
    rows, T = eng.rows(p), eng.frame_count(seg_len, fl, hop, True)
    units = eng.units_clips(n_units, seg_len, total_len=total, stride=seg_hop)

    # synthetic recording, resident in HBM (per-second blocks of sine + log-chirp + noise, amplitudes over 40 dB)
    y = torch.empty(total, dtype=torch.float32, device=dev)
    synth.torch_mixture_(y, sr, seed=1234 + rank, unit=sr)
    out = torch.empty((n_units, rows, T), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        eng.features_dev(y.data_ptr(), units, p, out.data_ptr(), stream)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    eng.profile_read(reset=True)
    eng.profile_enable(True)
    clocks = ClockSampler(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clk = clocks.stop()
    ms = e0.elapsed_time(e1)
    prof = eng.profile_read(reset=True)
    eng.profile_enable(False)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    unit_audio = w["seg_sec"] * (1.0 - w["overlap"])
    audio_s = n_units * unit_audio                       # unique audio seconds per rank per step
    value = world * audio_s * args.steps / (ms_max * 1e-3)

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D + kernels + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        yh = torch.empty(total, dtype=torch.float32, pin_memory=True)
        yh.copy_(y)
        oh = torch.empty((n_units, rows, T), dtype=torch.float32, pin_memory=True)
        oh_np = oh.numpy()
        torch.cuda.synchronize()
        eng.features_host(None, units, p, out=oh_np, y_ptr=yh.data_ptr())      # warm-up (allocates the lanes)
        ne = max(1, min(args.steps, args.e2e_steps))
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(ne):
            eng.features_host(None, units, p, out=oh_np, y_ptr=yh.data_ptr())
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        # the host path must reproduce the device path bit for bit
        same = bool(torch.equal(oh.to(dev), out))
        e2e = {"value": world * audio_s * ne / dt, "unit": "audio-s/s", "h2d_bytes_per_step": int(total * 4),
               "d2h_bytes_per_step": int(out.numel() * 4), "steps": ne, "ms_per_step": 1e3 * dt / ne,
               "matches_device_path": same, "host_placement": placement}
        del yh
        # additive ingest path: the same recording as 16-bit PCM (what the WAV files hold); the float32 line
        # above stays the headline because it is the reference's own in-memory format
        y16 = torch.empty(total, dtype=torch.int16, pin_memory=True)
        y16.copy_((y * 32767.0).round().clamp(-32768, 32767).to(torch.int16))
        torch.cuda.synchronize()
        eng.features_host_pcm16(y16.data_ptr(), units, p, oh_np)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(ne):
            eng.features_host_pcm16(y16.data_ptr(), units, p, oh_np)
        dt16 = time.perf_counter() - t0
        tt = torch.tensor([dt16], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt16 = float(tt.item())
        e2e["pcm16"] = {"value": world * audio_s * ne / dt16, "unit": "audio-s/s", "h2d_bytes_per_step": int(total * 2),
                        "d2h_bytes_per_step": int(out.numel() * 4), "ms_per_step": 1e3 * dt16 / ne}
        del y16, oh
    os.sched_setaffinity(0, full_affinity)                 # the CPU baseline below uses every host core again

    gather_ms = None
    if world > 1:
        g = torch.empty((world,) + tuple(out.shape), dtype=out.dtype, device=dev)
        torch.cuda.synchronize(); dist.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        dist.all_gather_into_tensor(g, out)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = g0.elapsed_time(g1)
        del g

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak, peak_src = (float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        alg_bytes = 4.0 * total + 4.0 * out.numel()                       # unique input samples + output floats, per step
        f_ms, f_n = prof["frame"]
        z_ms, z_n = prof["finalize"]
        # dominant kernel = the frame kernel; one launch processes one workspace chunk of units (the last chunk is shorter),
        # so per-launch bytes / per-launch time == per-step bytes / per-step kernel time
        launches_per_step = f_n / args.steps if args.steps else 0
        achieved = alg_bytes * args.steps / (f_ms * 1e-3) / 1e9 if f_ms > 0 else None
        traffic = None                                                    # DRAM bytes of one launch from the committed ncu --set full capture
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tr.get("workload") == args.workload:
                traffic = {"dram_bytes_per_launch": tr["dram_bytes_read"] + tr["dram_bytes_write"],
                           "algorithmic_bytes_of_that_launch": tr["algorithmic_bytes"], "source": tr["source"]}
        except Exception:
            pass
        line = {
            "metric": "audio-sec/sec (MFCC+spectral feats)", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {w['desc']}", "per_gpu_audio_hours": hours, "units_per_gpu": n_units,
                       "rows": rows, "frames_per_unit": T, "l2": "inputs larger than L2 (%.2f GB samples per step)" % (total * 4 / 1e9),
                       "parallelism": f"units sharded over {world} GPU(s), no data-path collective"},
            "roofline": {"bound": "hbm", "kernel": "frame_kernel (framing+window+rFFT+fused feature epilogues)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "peak_source": peak_src, "traffic": traffic,
                         "kernel_launches_per_step": launches_per_step,
                         "kernel_ms_per_launch": (f_ms / f_n) if f_n else None,
                         "algorithmic_bytes_per_launch": (alg_bytes / launches_per_step) if launches_per_step else None,
                         "kernel_ms_per_step": f_ms / args.steps, "kernel_share_of_step": f_ms / ms if ms > 0 else None,
                         "finalize_ms_per_step": z_ms / args.steps, "algorithmic_bytes_per_step": alg_bytes},
            "clocks": clk, "gpu_launches": int(f_n + z_n),
        }
        if e2e:
            line["e2e"] = e2e
        if gather_ms is not None:
            line["gather_ms"] = gather_ms
        if world == 1 and not args.no_cpu:
            # bounded CPU sample of the same workload, in a fresh process (fork pool; keeps CUDA out of the children)
            cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload,
                   "--steps", "1", "--warmup", "0", "--cpu-units", str(96 * len(os.sched_getaffinity(0)))]
            try:
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
                ref = json.loads(r.stdout.strip().splitlines()[-1])
                line["cpu_baseline"] = ref["cpu_baseline"]
            except Exception as exc:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": "audio-s/s", "cores": None, "kind": "port", "sample": f"failed: {exc}"}
        emit_json(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--hours", type=float, default=0.0, help="audio hours per GPU (default: the workload's)")
    ap.add_argument("--ws-mb", type=int, default=0, help="engine workspace limit in MiB (0: library default)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-units", type=int, default=0, help="segments per step of the CPU reference arm")
    args = ap.parse_args()
    guard_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
