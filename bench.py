#!/usr/bin/env python
"""bench.py — the batched calibration solve on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config cfg2|cfg3|cfg4|cfg5|ba]

A step = one pass of the hot path over one batch: stage the matches (davo_stage_matches) and solve
every problem (davo_solve_calibration); with N > 1 also the single all-gather of the solved records.
Workload at every N: BASELINE.json configs[1] per GPU — 65 536 problems x 256 matches, n = 10, float32,
error_threshold 1e-7 (the driver's eval value, networks/calibration_network.py:44), 1000-iteration cap —
i.e. weak scaling (N x 64K problems in total, each rank generating its own shard from seed 0xB200+rank).

`value` (solves/s) is device time with the raw inputs already in HBM; `e2e` is the same metric through
BFGSSolver.forward with HOST pinned buffers, H2D and D2H copies inside the timed region.
`--impl reference` times the CPU arm: the reference is pure Python and /root/reference does not exist on
the GPU box, so it is the C oracle port of the reference's algorithm (oracle/calib_oracle.c) on all host
threads, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# SURVEY.md §8(d); angle_ba: counted from csrc/objectives_ba.cuh per (view, point) pair per cost+gradient evaluation
FLOP_PER_MATCH_EVAL = {"distort10": 93.0, "joint": 145.0, "angle_ba": 200.0}
L2_BYTES = 126 << 20
CFG5_TOTAL = 1 << 20   # BASELINE.json configs[4]: 1M problems sharded over the GPUs (strong scaling)
SOLVER_KW = {"cfg2": dict(error_threshold=1e-7, iterations=1000),
             "cfg5": dict(error_threshold=1e-7, iterations=1000),
             "cfg3": dict(error_threshold=1e-5, iterations=1000),
             "cfg4": dict(error_threshold=1e-5, iterations=1000),
             # the entry script's objective with the driver's eval threshold
             # (networks/calibration_network.py:44), 4 views x 8 points as in its config
             "ba": dict(error_threshold=1e-7, iterations=1000)}


def make_batch(config: str, B: int, seed: int, dtype=np.float32):
    import davo_b200
    syn = davo_b200.synthetic
    if dtype != np.float32:  # --dtype f64: the entry script's own precision (float_precision="64"), side runs only
        gen = {"cfg2": lambda: syn.make_distort10(B, 256, seed=seed, dtype=dtype),
               "cfg4": lambda: syn.make_distort10(B, 256, seed=seed, dtype=dtype, ill_conditioned=True, pathological=0.02),
               "cfg3": lambda: syn.make_joint(B, 256, 4, seed=seed, dtype=dtype),
               "ba": lambda: syn.make_angle_ba(B, 8, 4, seed=seed, dtype=dtype)}
        return gen[config]()
    if config == "cfg2":
        return syn.make_distort10(B, 256, seed=seed, dtype=np.float32)
    if config == "cfg5":  # 1M problems: generated with torch on the GPU when there is one (numpy needs minutes)
        import torch
        return syn.make_distort10_torch(B, 256, seed=seed, device="cuda" if torch.cuda.is_available() else "cpu")
    if config == "cfg4":
        return syn.make_distort10(B, 256, seed=seed, dtype=np.float32, ill_conditioned=True, pathological=0.02)
    if config == "cfg3":
        return syn.make_joint(B, 256, 4, seed=seed, dtype=np.float32)
    if config == "ba":
        return syn.make_angle_ba(B, 8, 4, seed=seed, dtype=np.float32)
    raise ValueError(config)


def default_B(config: str, world: int = 1) -> int:
    if config == "cfg5":
        return CFG5_TOTAL // world
    return 16384 if config == "cfg3" else 65536


def workload_name(config: str, B: int) -> str:
    return {"cfg2": f"configs[1]: intrinsics+distortion fit, {B} problems x 256 matches, n=10",
            "cfg5": f"configs[4]: {CFG5_TOTAL} intrinsics+distortion problems x 256 matches sharded by problem, "
                    f"{B} per GPU, n=10",
            "cfg3": f"configs[2]: joint intrinsics + 4 view poses, {B} problems x 1024 matches, n=34",
            "cfg4": f"configs[3]: ill-conditioned heavy distortion, {B} problems x 256 matches, n=10",
            "ba": f"entry-script objective (SURVEY 8f row 1): bundle adjustment with angular error, {B} problems x "
                  f"4 views x 8 points, n=45"}[config]


def bench_config(config: str, B: int, solver_kw: dict) -> dict:
    """The `config` object of the JSON line: identical on the GPU arm and on `--impl reference`."""
    return {"workload": workload_name(config, B), "problems_per_gpu": B, "solver": solver_kw, "seed": "0xB200 + rank"}


def algorithmic_flops(batch, fevals: np.ndarray, iters: np.ndarray) -> float:
    """sum fevals * matches * F_fg + sum iters * (12 n^2 + 10 n), with the reference-equivalent
    evaluation count (what the reference's algorithm evaluates, not what the kernel skipped by reuse)."""
    matches = batch.N * batch.views
    n = batch.n
    return float(fevals.astype(np.float64).sum() * matches * FLOP_PER_MATCH_EVAL[batch.model]
                 + iters.astype(np.float64).sum() * (12 * n * n + 10 * n))


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md), through NVML
    (the library behind nvidia-smi) from a polling thread so that even a 100 ms region gets tens of samples."""

    BAD = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}

    def __init__(self, index: int):
        self.index, self.stop_flag, self.sm, self.reasons, self.thread, self.max_mhz = index, False, [], set(), None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = int(vis.split(",")[self.index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.reasons.add(f"nvml unavailable: {type(e).__name__}")
            return

        def poll():
            while not self.stop_flag:
                try:
                    self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                    mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                    for name, bit in self.BAD.items():
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:  # noqa: BLE001
                    pass
                time.sleep(0.004)

        self.thread = threading.Thread(target=poll, daemon=True)
        self.thread.start()

    def stop(self) -> dict:
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "samples": len(self.sm), "reasons": sorted(self.reasons)}


def cpu_arm(config: str, B_workload: int, budget_s: float, threads: int = 0):
    """Time the C oracle (a port of the reference's algorithm) on a bounded sample of the workload."""
    from oracle import c_oracle
    # torchrun exports OMP_NUM_THREADS=1: ask for every core this process may run on explicitly
    cores = len(os.sched_getaffinity(0)) if threads <= 0 else threads
    kw = SOLVER_KW[config]
    pilot_B = max(16 * cores, 256)
    pilot = make_batch(config, pilot_B, 0xB200)
    t0 = time.perf_counter()
    c_oracle.solve_batch(pilot, threads=cores, **kw)
    pilot_t = max(time.perf_counter() - t0, 1e-4)
    sample_B = int(min(B_workload, max(pilot_B, budget_s / pilot_t * pilot_B)))
    sample = make_batch(config, sample_B, 0xB200)
    t0 = time.perf_counter()
    r = c_oracle.solve_batch(sample, threads=cores, **kw)
    dt = time.perf_counter() - t0
    return dict(value=sample_B / dt, seconds=dt, cores=cores, sample_B=sample_B,
                iters_per_s=float(r["iters"].sum()) / dt,
                sample=f"first {sample_B} problems of the workload (seed 0xB200), {dt:.1f} s on {cores} threads")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # N > 1: rank 0 alone runs and prints it
    B = args.batch or default_B(args.config)
    steps, warm = max(args.steps, 1), args.warmup
    per_step_budget = min(20.0, 120.0 / (steps + warm))
    for _ in range(warm):
        cpu_arm(args.config, B, per_step_budget)
    res = [cpu_arm(args.config, B, per_step_budget) for _ in range(steps)]
    value = float(np.mean([r["value"] for r in res]))
    ms = float(np.mean([r["seconds"] for r in res])) * 1e3
    line = {"impl": "reference", "metric": "calibration_solves_per_sec", "value": value, "unit": "solves/s",
            "bfgs_iters_per_sec": float(np.mean([r["iters_per_s"] for r in res])),
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args.config, B, SOLVER_KW[args.config]),
            "cpu_baseline": {"value": value, "unit": "solves/s", "cores": res[-1]["cores"], "kind": "port",
                             "sample": res[-1]["sample"]},
            "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference is pure Python/PyTorch and is not present on the GPU box; this arm times the C "
                    "port of its algorithm (oracle/calib_oracle.c, -O2, OpenMP over problems)"}
    emit(line)


_JSON_FD = None


def _reserve_stdout():
    """Keep the real stdout for the ONE JSON line and point fd 1 at stderr for everything else: NCCL (and any
    other native library) prints banners such as "NCCL version ..." straight to fd 1."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    _reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg3", "cfg4", "cfg5", "ba"])
    ap.add_argument("--batch", type=int, default=0, help="problems per GPU (default: the BASELINE size)")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"],
                    help="arithmetic type (BASELINE asks for f32; the entry script itself trains in f64)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--in-flight", type=int, default=0,
                    help="batches in flight in the throughput pass (streams); 0 = 3 on one GPU, 2 with the per-step "
                         "all-gather (measured at 8 GPUs: 3.86 ms per step with 2 in flight, 3.93 with 3)")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="time the K steps one batch at a time on one stream (no second batch in flight)")
    ap.add_argument("--no-side-configs", action="store_true",
                    help="skip the 'configs' object (cfg3, cfg4, bundle adjustment, cfg5 strong scaling)")
    ap.add_argument("--side-steps", type=int, default=12, help="timed steps per side configuration")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    import davo_b200
    from davo_b200 import _lib
    from davo_b200.distributed import ResultSlab

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # N ranks stream host buffers over N PCIe links at once: keep each rank (and the pinned memory it is about
        # to allocate) on the CPUs / memory node next to its GPU
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = int(vis.split(",")[local]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else local
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(phys))
        except Exception:  # noqa: BLE001 (no NVML / no topology information in this VM: run unpinned)
            pass
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    props = torch.cuda.get_device_properties(dev)
    try:
        traffic_json = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except OSError:
        traffic_json = {}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def measure(config: str, B: int, dtype: str, steps: int, warm: int, solver_kw: dict, with_e2e: bool,
                with_clocks: bool):
        """One workload: W warm-up steps, K timed steps (stage + solve (+ all-gather)), device-timed, max over ranks.
        Returns the fields of the JSON line that describe this workload."""
        np_dt = np.float64 if dtype == "f64" else np.float32
        t_dt = torch.float64 if dtype == "f64" else torch.float32
        esz = 8 if dtype == "f64" else 4
        batch = make_batch(config, B, 0xB200 + rank, np_dt)
        solver = davo_b200.BFGSSolver(**solver_kw).eval()

        # ---- raw inputs resident in HBM (value) and in pinned host memory (e2e) ------------------------------
        # (angle_ba has no 3-D points: its second input is the visibility mask)
        second = batch.weights if batch.model == "angle_ba" else batch.points_3d
        h_pts, h_obs, h_x0 = (torch.from_numpy(a) for a in (second, batch.obs, batch.x0))
        if with_e2e:
            h_pts, h_obs, h_x0 = h_pts.pin_memory(), h_obs.pin_memory(), h_x0.pin_memory()
        d_pts, d_obs, d_x0 = h_pts.to(dev), h_obs.to(dev), h_x0.to(dev)
        input_bytes = (h_pts.numel() + h_obs.numel() + h_x0.numel()) * esz
        # timing rule: inputs larger than L2, or flush L2 between timed iterations (outside the per-step events)
        flush = torch.empty(2 * L2_BYTES, dtype=torch.uint8, device=dev) if input_bytes <= L2_BYTES else None
        slab = ResultSlab(B * world, batch.n, t_dt, world, dev)
        out = slab.buffers(rank)
        # Throughput pass: the K timed steps alternate between two streams, each with its own result slab (a step
        # allocates its own staged matches), so the tail of one launch — the last problems of a persistent grid, or
        # config 4's stragglers on a handful of SMs — overlaps the head of the next.  Only without the L2 flush
        # (its per-step events need the steps one after the other).
        pipelined = flush is None and not args.no_pipeline
        n_fly = max(2, int(args.in_flight)) if args.in_flight else (3 if world == 1 else 2)
        slabs = [slab] + [ResultSlab(B * world, batch.n, t_dt, world, dev) for _ in range(n_fly - 1)] if pipelined else [slab]
        streams = [torch.cuda.Stream(device=dev) for _ in range(n_fly)] if pipelined else None

        def make_objective(pts, obs):
            if batch.model == "distort10":
                return davo_b200.DistortionObjective(pts, obs)  # runs davo_stage_matches
            if batch.model == "angle_ba":
                return davo_b200.AngleDistanceObjective(obs, pts)
            return davo_b200.JointPoseObjective(pts, obs)

        def step(ev=None, which=0):
            sl = slabs[which]
            if flush is not None:
                flush.zero_()
            if ev:
                ev[0].record()
            obj = make_objective(d_pts, d_obs)
            if ev:
                ev[1].record()
            solver.solve_into(d_x0, obj, out=sl.buffers(rank))
            if ev:
                ev[2].record()
            sl.all_gather(rank)
            if ev:
                ev[3].record()

        for _ in range(warm):
            step()
        barrier()
        sampler = ClockSampler(local)
        if rank == 0 and with_clocks:
            sampler.start()
        launches0 = _lib.launch_count()
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        events = []
        barrier()
        t_start.record()
        for _ in range(steps):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            step(ev)  # per-kernel events are read after the loop; recording them does not synchronise
            events.append(ev)
        t_end.record()
        barrier()
        launches = _lib.launch_count() - launches0
        clocks = sampler.stop() if (rank == 0 and with_clocks) else None
        # with an L2 flush between the steps the timed region is the sum of the per-step event pairs (flush excluded)
        total_ms = (t_start.elapsed_time(t_end) if flush is None
                    else float(sum(e[0].elapsed_time(e[3]) for e in events)))
        solve_kernel_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in events]))
        stage_kernel_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in events]))
        sequential_ms = total_ms
        if pipelined:
            main_stream = torch.cuda.current_stream()
            for st in streams:                       # warm-up on the two streams themselves: their allocator pools
                st.wait_stream(main_stream)          # and first launches are paid before the timed region
            for k in range(n_fly * max(warm, 2)):
                with torch.cuda.stream(streams[k % n_fly]):
                    step(None, k % n_fly)
            for st in streams:
                main_stream.wait_stream(st)
            barrier()
            launches0 = _lib.launch_count()
            t_start.record()
            for st in streams:
                st.wait_stream(main_stream)
            for k in range(steps):
                with torch.cuda.stream(streams[k % n_fly]):
                    step(None, k % n_fly)
            for st in streams:
                main_stream.wait_stream(st)
            t_end.record()
            barrier()
            launches = _lib.launch_count() - launches0
            total_ms = t_start.elapsed_time(t_end)
        pipelined_ms = total_ms if pipelined else None
        if world > 1:
            t = torch.tensor([total_ms, sequential_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms, sequential_ms = float(t[0].item()), float(t[1].item())
            pipelined_ms = total_ms if pipelined else None
        # both passes time the same K steps the same way; `value` is the schedule with the higher throughput (two
        # batches in flight lose when a batch is large enough that the next batch's staging kernel, at the HBM rate,
        # delays the running solve's bulk loads more than the overlapped tail gains: config 5 at 512K problems per GPU)
        used = "pipelined" if (pipelined and total_ms <= sequential_ms) else "sequential"
        if used == "sequential":
            total_ms = sequential_ms
        ms_per_step = total_ms / steps

        iters = out.iterations.cpu().numpy()
        fevals = out.evaluations.cpu().numpy()
        converged = float(out.converged.float().mean().item())
        reasons = np.bincount(out.reason.cpu().numpy(), minlength=4).tolist()
        stats = torch.tensor([float(iters.sum()), float(fevals.sum()), converged * B], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(stats)
        total_B = B * world
        value = total_B / (ms_per_step * 1e-3)

        # ---- e2e: the public call with host buffers, H2D + D2H inside the timed region -----------------------
        e2e = None
        if with_e2e:
            def e2e_step():
                obj = make_objective(h_pts, h_obs)                       # H2D of points + observations, staging
                return solver(h_x0, obj, return_info=True)               # H2D of x0, solve, D2H of every output
            for _ in range(3):
                e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                info = e2e_step()
            barrier()
            e2e_s = (time.perf_counter() - t0) / steps
            if world > 1:
                t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                e2e_s = float(t.item())
            d2h = sum(t.numel() * t.element_size() for t in info)
            e2e = {"value": total_B / e2e_s, "unit": "solves/s", "h2d_bytes_per_step": int(input_bytes),
                   "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3,
                   "api": "BFGSSolver.forward(host tensors, DistortionObjective(host tensors), return_info=True), one batch "
                          "at a time (submitting a second batch while the first finishes was tried: 6.4 to 13 ms per "
                          "batch from run to run, so the blocking call is what is timed)",
                   "h2d_gbs_per_gpu": input_bytes / e2e_s / 1e9,
                   "note": "bound by the host-to-device copy of the raw inputs (tools/h2d_probe.py measures the "
                           "plain-copy ceiling of the same bytes; DESIGN.md section 6)"}

        # ---- roofline of the dominant kernel (the solve kernel): FP32 / FP64 CUDA-core pipe ------------------
        sm_max_mhz = float(peaks.get("sm_max_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0)
        lanes = 128 if dtype == "f32" else 64   # FP32 / FP64 CUDA-core lanes per SM on B200
        pipe_peak = props.multi_processor_count * lanes * 2 * sm_max_mhz * 1e6 / 1e12   # TFLOP/s
        flops = algorithmic_flops(batch, fevals, iters)                                 # this rank's launch
        achieved = flops / (solve_kernel_ms * 1e-3) / 1e12
        stage_bytes = batch.N * B * (20 + 16) * (esz // 4) if batch.model == "distort10" else 0
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        stage_gbs = stage_bytes / (stage_kernel_ms * 1e-3) / 1e9 if stage_bytes else None
        traffic = stage_traffic = None
        try:  # DRAM bytes per launch of the solve kernel from the committed ncu --set full capture of this workload
            if B == default_B(config, world) and config in ("cfg2", "cfg3", "cfg4") and dtype == "f32":
                traffic = traffic_json["solve_" + config]["bytes"]
            if config == "ba" and dtype == "f32":   # captured at 32K problems: scale to this launch
                traffic = int(traffic_json["solve_ba"]["bytes"] * B / traffic_json["solve_ba"]["problems"])
            stage_traffic = traffic_json["stage_cfg2"]["bytes"] if (config == "cfg2" and B == default_B("cfg2")) else None
        except KeyError:
            pass
        res = {
            "value": value, "unit": "solves/s",
            "bfgs_iters_per_sec": float(stats[0].item()) / (ms_per_step * 1e-3),
            "fevals_per_sec": float(stats[1].item()) / (ms_per_step * 1e-3),
            "converged_frac": float(stats[2].item()) / total_B,
            "steps": steps, "warmup": warm, "ms_per_step": ms_per_step, "dtype": dtype,
            "config": bench_config(config, B, solver_kw),
            "timing": {"l2": (f"inputs larger than L2 ({input_bytes / 1e6:.0f} MB of raw inputs per GPU)" if flush is None
                              else f"L2 flushed between timed steps (a {2 * L2_BYTES >> 20} MiB buffer is rewritten, "
                                   f"outside the per-step events; inputs are {input_bytes / 1e6:.0f} MB per GPU)"),
                       "collective": "one all_gather_into_tensor of the solved records"
                       if world > 1 else "none (single GPU)",
                       "pipeline": (f"the K timed steps alternate between {n_fly} CUDA streams ({n_fly} batches in flight): the tail of "
                                    "one launch overlaps the head of the next; ms_per_step_sequential is the same K steps "
                                    "one batch at a time on one stream; value / ms_per_step are the faster schedule: "
                                    + used) if pipelined else "one batch at a time"},
            "ms_per_step_sequential": sequential_ms / steps,
            "ms_per_step_pipelined": None if pipelined_ms is None else pipelined_ms / steps,
            "kernel_ms": {"solve": solve_kernel_ms, "stage": stage_kernel_ms},
            "mean_iters": float(iters.mean()), "mean_fevals": float(fevals.mean()), "reasons_rank0": reasons,
            "roofline": {"bound": "fp32" if dtype == "f32" else "fp64", "achieved": achieved, "peak": pipe_peak,
                         "unit": "TFLOP/s", "frac": achieved / pipe_peak, "traffic": traffic,
                         "traffic_note": "DRAM bytes per launch (ncu); the kernel is FP32-pipe bound, its per-problem "
                                         "inputs are read from HBM once and kept in shared memory",
                         # evaluations the kernel really ran = reference-equivalent evaluations - reused probes
                         "executed_frac": achieved / pipe_peak * float((fevals - iters).sum()) / max(float(fevals.sum()), 1.0),
                         "peak_source": f"{props.multi_processor_count} SMs x {lanes} lanes x 2 x {sm_max_mhz:.0f} MHz "
                                        "(no FP32 figure in MEASURED_PEAKS.json; sm_max_mhz taken from it)",
                         "flops_counted": "SURVEY.md 8(d): reference-equivalent fevals x matches x "
                                          f"{FLOP_PER_MATCH_EVAL[batch.model]:.0f} + iters x (12n^2+10n)"},
            "roofline_staging": None if stage_gbs is None else {
                "bound": "hbm", "achieved": stage_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": stage_gbs / hbm_peak,
                "traffic": stage_traffic, "bytes_counted": "20 B read + 16 B written per match"},
            "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e,
        }
        del d_pts, d_obs, d_x0, slab, slabs, out, flush
        torch.cuda.empty_cache()
        return res

    B = args.batch or default_B(args.config, world)
    kw = SOLVER_KW[args.config]
    steps, warm = args.steps, max(args.warmup, 3)
    head = measure(args.config, B, args.dtype, steps, warm, kw, with_e2e=not args.no_e2e, with_clocks=True)

    # ---- the other BASELINE configurations, on the record beside the headline (fewer steps each) ---------------
    side = {}
    if not args.no_side_configs and args.config == "cfg2" and not args.batch:
        side_steps = max(3, min(steps, args.side_steps))
        plan = [("cfg5_strong", "cfg5", "f32", SOLVER_KW["cfg5"])]
        if world == 1:
            plan = [("cfg3", "cfg3", "f32", SOLVER_KW["cfg3"]), ("cfg4", "cfg4", "f32", SOLVER_KW["cfg4"]),
                    ("ba_f32_thr1e-4", "ba", "f32", dict(error_threshold=1e-4, iterations=1000)),
                    ("ba_f64_thr1e-7", "ba", "f64", SOLVER_KW["ba"])] + plan
        for label, cfg, dt, skw in plan:
            r = measure(cfg, default_B(cfg, world), dt, side_steps, 3, skw, with_e2e=False, with_clocks=False)
            side[label] = {"workload": r["config"]["workload"], "solver": skw, "dtype": dt, "value": r["value"],
                           "unit": "solves/s", "ms_per_step": r["ms_per_step"],
                           "ms_per_step_sequential": r["ms_per_step_sequential"],
                           "ms_per_step_pipelined": r["ms_per_step_pipelined"], "pipeline": r["timing"]["pipeline"],
                           "steps": side_steps,
                           "kernel_ms": r["kernel_ms"]["solve"], "bfgs_iters_per_sec": r["bfgs_iters_per_sec"],
                           "mean_iters": r["mean_iters"], "mean_fevals": r["mean_fevals"],
                           "converged_frac": r["converged_frac"], "reasons_rank0": r["reasons_rank0"],
                           "roofline": {"frac": r["roofline"]["frac"], "executed_frac": r["roofline"]["executed_frac"],
                                        "achieved": r["roofline"]["achieved"], "peak": r["roofline"]["peak"],
                                        "unit": "TFLOP/s", "bound": r["roofline"]["bound"]},
                           "scaling": "strong" if cfg == "cfg5" else "weak", "n_gpus": world,
                           "l2": r["timing"]["l2"]}

    # ---- the fused initial-guess network: one tcgen05 kernel in front of the solve -------------
    if side and world == 1:
        net = davo_b200.CalibrationNetwork(4, 8).to(dev).eval()
        rows = 65536
        xin = torch.randn(rows, 64, device=dev)

        def timed(fn, reps=20):
            for _ in range(3):
                fn(xin)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn(xin)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        with torch.no_grad():
            fused_ms = timed(net.estimate)
            modules_ms = timed(net.initial_estimator)
        flop = rows * 2.0 * (64 * 256 + 256 * 256 + 256 * 45)
        tensor_peak = float(peaks.get("bf16_tflops", 1590.0)) / 2.0   # dense TF32 is half the bf16 rate
        side["initial_guess_mlp"] = {
            "workload": "CalibrationNetwork.initial_estimator in eval mode, 65536 rows, 64-256-256-45 "
                        "(Linear-GELU-BatchNorm x2 + Linear), float32",
            "value": rows / (fused_ms * 1e-3), "unit": "rows/s", "ms_per_step": fused_ms,
            "torch_modules_ms": modules_ms,
            "roofline": {"bound": "tensor", "achieved": 3.0 * flop / (fused_ms * 1e-3) / 1e12, "peak": tensor_peak,
                         "unit": "TFLOP/s", "frac": 3.0 * flop / (fused_ms * 1e-3) / 1e12 / tensor_peak,
                         "note": "TF32 MMAs executed (3 per float32 product: hi/lo operand split); peak = measured "
                                 "bf16 burst rate / 2; float32-equivalent rate = achieved / 3"}}
        del net, xin

    # ---- BASELINE configs[0]: the reference's own CPU-sized case, one call on 64 problems (latency, not throughput) --
    if side and world == 1 and rank == 0:
        import time as _time
        small = make_batch("cfg2", 64, 0xB200)
        s_obj = davo_b200.DistortionObjective(torch.from_numpy(small.points_3d).to(dev), torch.from_numpy(small.obs).to(dev))
        s_x0 = torch.from_numpy(small.x0).to(dev)
        s_solver = davo_b200.BFGSSolver(**SOLVER_KW["cfg2"]).eval()
        calls = []
        for rep in range(23):
            torch.cuda.synchronize()
            t0 = _time.perf_counter()
            s_info = s_solver(s_x0, s_obj, return_info=True)
            torch.cuda.synchronize()
            calls.append(_time.perf_counter() - t0)
        call_ms = 1e3 * float(np.median(calls[3:]))
        from oracle import c_oracle   # the checker, timed as the CPU baseline of this case (cpu_baseline leg)
        t0 = _time.perf_counter()
        c_oracle.solve_batch(small, **SOLVER_KW["cfg2"])
        port_ms = 1e3 * (_time.perf_counter() - t0)
        side["cfg1_latency"] = {
            "workload": "configs[0]: the reference's CPU-sized case, ONE call on 64 problems x 256 matches, n=10, float32 "
                        "(device-resident inputs, host clock around BFGSSolver.forward + synchronize, median of 20)",
            "solver": SOLVER_KW["cfg2"], "dtype": "f32", "ms_per_call": call_ms, "value": 64.0 / (call_ms * 1e-3),
            "unit": "solves/s", "max_fevals": int(s_info.evaluations.max()),
            "cpu_port_ms": port_ms, "cpu_port_cores": os.cpu_count(),
            "note": "latency bound: one warp per problem, the call lasts as long as its slowest problem"}
        del s_obj, s_x0

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    line = {"metric": "calibration_solves_per_sec", "value": head["value"], "unit": "solves/s",
            "bfgs_iters_per_sec": head["bfgs_iters_per_sec"], "fevals_per_sec": head["fevals_per_sec"],
            "converged_frac": head["converged_frac"], "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if args.config == "cfg5" else "weak", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic"}
    for k in ("config", "timing", "ms_per_step_sequential", "ms_per_step_pipelined", "kernel_ms", "mean_iters", "mean_fevals", "reasons_rank0", "roofline", "roofline_staging",
              "clocks", "gpu_launches", "e2e"):
        line[k] = head[k]
    if side:
        line["configs"] = side
    if not args.no_cpu_baseline and world == 1:
        c = cpu_arm(args.config, B, 15.0)
        line["cpu_baseline"] = {"value": c["value"], "unit": "solves/s", "cores": c["cores"], "kind": "port",
                                "sample": c["sample"]}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
