#!/usr/bin/env python
"""bench.py - registered frame-pairs/sec for shoot (S=10 EPDiff steps) + warp + 126-sector strain.

Contract (see the task statement / DESIGN.md section "Measurement"):
  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

Workload = BASELINE.json configs[1]: batch of 64 slices x 25 frames, 128x128, per GPU
(P = 1536 frame-pairs per GPU per step, weak scaling: slices shard across ranks with no
data-path collective).  A "step" is one pass of the hot path over that batch.

* value    : whole-job frame-pairs/s with inputs resident in HBM (CUDA events, max over ranks)
* e2e      : same metric through the public API with HOST (pinned) inputs: H2D of v0 + masks and
             D2H of the strain matrices inside the timed region
* roofline : algorithmic bytes (700*N per pair) / duration of the fused
             shooting kernel, against MEASURED_PEAKS.json hbm_gbs
* cpu_baseline : the plain-C OpenMP oracle (a port: the reference's own lagomorph path is not runnable)
             on a bounded sample, rank 0, N=1 only
* extra    : the other BASELINE.json configs measured in the same run on the same ranks (short step counts):
             c3_train (configs[2]), c4_sharded (configs[3], strong-sharded over the N ranks), c5_step (configs[4],
             full training step with the NCCL gradient all-reduce)
`--impl reference` times that same CPU oracle with all host threads (the reference's CPU path); the thread count
is passed explicitly, so torchrun's OMP_NUM_THREADS=1 does not shrink it.
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib
import statistics
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

METRIC = "registered frame-pairs/sec (shoot+warp+strain, 128^2)"
UNIT = "frame-pairs/s"
B_PER_GPU, T_FRAMES, H, W, S_STEPS = 64, 25, 128, 128, 10
PARAMS = (1.0, 0.1, 0.05)
N_SECTORS, N_FRAMES = 126, 40
BYTES_PER_PAIR = 4 * (15 + 16 * S_STEPS) * H * W          # 700*N at S=10


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample-slices", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the configs[2]/[3]/[4] measurements")
    ap.add_argument("--mask-format", default="bits", choices=["bits", "u8"],
                    help="host format of the binary cine masks on the e2e leg: numpy.packbits (1 bit/pixel) or uint8")
    return ap.parse_args()


def workload_config(n_gpus):
    return {"workload": f"configs[1]: {B_PER_GPU} slices x {T_FRAMES} frames {H}x{W} per GPU, "
                        f"P={B_PER_GPU * (T_FRAMES - 1)} frame-pairs/GPU/step, EPDiff S={S_STEPS}, "
                        f"forward shooting + warp + {N_SECTORS}-sector strain",
            "slices_per_gpu": B_PER_GPU, "frames": T_FRAMES, "grid": [H, W], "epdiff_steps": S_STEPS,
            "fluid_params": list(PARAMS), "parallelism": f"slice-sharded x{n_gpus}, no data-path collective",
            "l2_policy": "inputs larger than L2 (v0 201 MB + masks 105 MB per step vs 126 MB L2)"}


def kernel_source_digest():
    """sha256 over the sources the dominant kernel is compiled from: stamps profiles/traffic.json."""
    import hashlib
    csrc = next(ROOT.glob("*_b200")) / "csrc"
    h = hashlib.sha256()
    for name in ("shoot.cu", "fft.cuh", "common.cuh", "strain.cuh"):
        h.update((csrc / name).read_bytes())
    return h.hexdigest()


def measured_traffic():
    """dram__bytes_read+write per launch of the fused kernel from the committed ncu capture (profiles/).  The capture
    is stamped with the digest of the kernel's sources: when they have changed since, the number is stale -> None."""
    p = ROOT / "profiles" / "traffic.json"
    try:
        rec = json.loads(p.read_text())
        if rec.get("kernel_source_sha256") != kernel_source_digest():
            return None
        return int(rec["traffic_bytes_per_launch"])
    except Exception:
        return None


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock + throttle reasons during the timed region (nvidia-smi recipe line)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.005)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def bind_to_gpu_numa_node(index):
    """Pin this rank to the CPUs local to its GPU before any pinned host memory is allocated, so the staging
    buffers are first-touched on the GPU's NUMA node (matters for the host->device leg at N >= 4)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def make_inputs(pkg, n_slices, seed):
    vol = pkg.synthetic.synthetic_masks(n_slices, T_FRAMES, H, W, seed=seed)                  # (B,1,T,H,W)
    v0 = pkg.synthetic.synthetic_v0(n_slices * (T_FRAMES - 1), H, W, seed=seed + 1, max_disp=3.0)
    return vol, v0


def cpu_forward(pkg, n_slices):
    """Return (callable running one pass of the CPU oracle on `n_slices` slices, description, threads).

    Preferred: the plain-C OpenMP oracle (oracle/lddmm_c.c) on all host threads; fallback: the torch-CPU oracle.
    Both are ports: the reference's own lagomorph CPU path is not runnable."""
    import oracle
    try:                      # the host threads this process may run on; NOT OMP_NUM_THREADS (torchrun sets it to 1)
        threads = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        threads = os.cpu_count() or 1
    vol, v0 = make_inputs(pkg, n_slices, seed=2434)
    try:
        from oracle import c_oracle
        c_oracle.lib()

        def run():
            return c_oracle.forward_volume(v0, vol, PARAMS, S_STEPS, N_SECTORS, N_FRAMES, nthreads=threads)
        return run, "plain-C OpenMP oracle (oracle/lddmm_c.c)", threads
    except Exception:
        torch.set_num_threads(threads)
        src_vol, tar_vol = oracle.split_vol_to_registration_pairs(vol, "Lagrangian", 3)
        metric = oracle.FluidMetric(PARAMS)

        def run():
            with torch.no_grad():
                return oracle.forward_volume(v0, src_vol, tar_vol, metric, S_STEPS, N_SECTORS, N_FRAMES)
        return run, "torch-CPU fp32 oracle", threads


def cpu_pairs_per_s(pkg, n_slices, reps):
    """The CPU oracle on a bounded sample of the same workload (checker used as baseline only)."""
    run, what, threads = cpu_forward(pkg, n_slices)
    times = []
    for i in range(reps + 1):
        t0 = time.perf_counter()
        run()
        if i > 0:
            times.append(time.perf_counter() - t0)
    return n_slices * (T_FRAMES - 1) / statistics.median(times), what, threads


def run_reference(args):
    """Reference arm: the CPU implementation of the path (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as g
    pkg = g.load_package()
    n_slices = args.cpu_sample_slices
    run, what, threads = cpu_forward(pkg, n_slices)
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = time.perf_counter() - t0
    pairs = n_slices * (T_FRAMES - 1)
    value = pairs * args.steps / dt
    sample = (f"each step = {n_slices} slices x {T_FRAMES - 1} pairs ({pairs} frame-pairs) of the configs[1] workload, "
              f"{what}, {threads} threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference's own lagomorph CPU path is not runnable (not vendored; SURVEY.md section 0): "
                    "this is the restated oracle port"}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import __graft_entry__ as g
    pkg = g.load_package()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs CUDA devices (the product path has no CPU fallback)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    try:
        all_cpus = os.sched_getaffinity(0)
    except (AttributeError, OSError):
        all_cpus = None
    bind_to_gpu_numa_node(local)         # pinned staging memory is first-touched on the GPU's NUMA node
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    lib, ptr, stream = pkg._lib.lib(), pkg._lib.ptr, pkg._lib.stream
    metric = pkg.FluidMetric(PARAMS)
    B, T1 = B_PER_GPU, T_FRAMES - 1
    P = B * T1

    # ---- inputs: this rank's shard of slices (weak scaling: B_PER_GPU slices per rank)
    vol_h, v0_h = make_inputs(pkg, B, seed=2434 + 17 * rank)
    vol_h, v0_h = vol_h.pin_memory(), v0_h.pin_memory()
    vol_d, v0_d = vol_h.to(dev), v0_h.to(dev)
    src_vol, tar_vol = pkg.data.split_vol_to_registration_pairs(vol_d, "Lagrangian", 3)   # views of vol_d

    def step_resident():
        with torch.no_grad():
            return pkg.shoot_warp_strain(v0_d, src_vol, tar_vol, metric, num_steps=S_STEPS,
                                         n_sectors=N_SECTORS, n_frames=N_FRAMES)

    d2h = B * N_SECTORS * N_FRAMES * 4

    # public host-buffer API: pinned host inputs -> strain matrices on the host; H2D of slice chunks on a
    # copy stream overlaps the fused kernel of the previous chunk.  The cine masks are binary (README.md:21) and are
    # handed over the way a dataset would store them, one BIT per pixel (numpy.packbits along the row; widened to the
    # fp32 staging volume on the device by b2_unpack_bits, no host pass inside or outside the timed region per step),
    # v0 as fp32.  --mask-format u8 selects the one-byte-per-pixel route of the earlier rounds.
    if args.mask_format == "bits":
        import numpy as np
        vol_u8_h = torch.from_numpy(np.packbits(vol_h.numpy() > 0.5, axis=-1)).pin_memory()
    else:
        vol_u8_h = vol_h.to(torch.uint8).pin_memory()
    pipe = pkg.HostPipeline(B, T_FRAMES, H, W, metric, num_steps=S_STEPS, n_sectors=N_SECTORS, n_frames=N_FRAMES,
                            device=dev)          # default chunking: four equal chunks of 16 slices
    pending = []

    def step_e2e():
        # streaming use of the public API: submit step i, then wait for step i-1's strain matrices on the host
        # (every step's result is host-visible inside the timed region; uploads of step i overlap kernels of i-1)
        pending.append(pipe.submit(v0_h, vol_u8_h))
        if len(pending) > 1:
            pending.pop(0).get()

    def drain_e2e():
        while pending:
            pending.pop(0).get()

    def timed(fn, steps, warmup, after=None):
        for _ in range(warmup):
            fn()
        if after:
            after()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = pkg._lib.launches()
        e0.record()
        for _ in range(steps):
            fn()
        if after:
            after()                                     # host sync on the last result, then the closing event
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = pkg._lib.launches() - l0
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches

    # bring the SM clocks out of idle with unrelated work (NOT the hot path): a cold GPU needs tens of
    # milliseconds to boost, which is longer than W short warm-up steps
    spin = torch.randn(4096, 4096, device=dev)
    t_spin = time.perf_counter()
    while time.perf_counter() - t_spin < 1.0:
        spin = (spin @ spin).clamp_(-1, 1)
        torch.cuda.synchronize()
    del spin

    sampler = ClockSampler(local)
    sampler.start()
    ms_total, launches = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop()
    ms_e2e, _ = timed(step_e2e, args.steps, max(args.warmup, 3), after=drain_e2e)
    h2d = int(pipe.h2d_bytes)                            # v0 as fp32 + binary masks (bits or bytes)

    # ---- H2D ceiling of this box at N ranks: the SAME pinned host buffers, the same bytes per rank per step, as two
    # plain cudaMemcpyAsync (v0, masks) with nothing else running, all ranks copying at once - what the host -> device
    # leg alone allows
    ceil_v0 = torch.empty_like(v0_d)
    ceil_vol = torch.empty(vol_u8_h.shape, dtype=torch.uint8, device=dev)

    def step_copy():
        ceil_v0.copy_(v0_h, non_blocking=True)
        ceil_vol.copy_(vol_u8_h, non_blocking=True)

    # best of three timed blocks: the ceiling is an upper bound, a perturbed block (seen once: 53.3 instead of 55.4 GB/s,
    # below what the pipeline itself moved in the same run) must not lower it
    ms_copy = min(timed(step_copy, args.steps, 3)[0] for _ in range(3))
    del ceil_v0, ceil_vol

    # ---- roofline of the dominant kernel: direct C-ABI launches of the fused shooting kernel
    out = step_resident()
    mom = pkg.strain.mask_moments(src_vol[:, 0, 0].contiguous())
    frame = pkg.strain.Frame(N_SECTORS, B, dev)
    fs = frame.c_struct()
    a = pkg._lib.ShootArgs()
    tar_flat = tar_vol.reshape(P, 1, H, W).contiguous()
    src0 = src_vol[:, :, 0].contiguous()
    counts = torch.empty((B, N_SECTORS, T1), dtype=torch.int32, device=dev)
    a.v0, a.src, a.tar, a.moments, a.table = v0_d.data_ptr(), src0.data_ptr(), tar_flat.data_ptr(), mom.data_ptr(), fs.table
    a.table_slice_stride, a.theta0, a.clockwise, a.flags = fs.table_slice_stride, fs.theta0, fs.clockwise, 0
    a.m0, a.vel, a.u = out["momentum"].data_ptr(), out["velocity"].data_ptr(), out["displacement"].data_ptr()
    a.sdef, a.S, a.counts, a.traj = out["deformed_source"].data_ptr(), out["strain_matrix"].data_ptr(), counts.data_ptr(), None
    a.B, a.T1, a.H, a.W = B, T1, H, W
    a.num_steps, a.src_per_pair, a.v0_is_momentum = S_STEPS, 0, 0
    a.n_sectors, a.n_frames, a.background = N_SECTORS, N_FRAMES, 0
    a.alpha, a.beta, a.gamma, a.T = PARAMS[0], PARAMS[1], PARAMS[2], 1.0
    nws = lib.b2_shoot_workspace_bytes(B, T1, H, W, S_STEPS)
    ws = torch.empty(nws, dtype=torch.uint8, device=dev)
    import ctypes

    def kernel_only():
        pkg._lib.check(lib.b2_shoot_fwd(ctypes.byref(a), ptr(ws), nws, stream()), "b2_shoot_fwd")

    def timed_launches(fn, steps, warmup):
        # average duration of the individual launches: one event pair around EACH launch on the launching stream, so
        # that a descheduled host thread between two launches (seen once: +2.8 % on the 20-launch region) is not
        # booked as kernel time; max over ranks like every other number
        for _ in range(warmup):
            fn()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a0, a1 in ev:
            a0.record()
            fn()
            a1.record()
        barrier()
        per = sorted(a0.elapsed_time(a1) for a0, a1 in ev)
        t = torch.tensor([sum(per) / len(per)], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), per[0], per[-1]

    k_ms, k_min, k_max = timed_launches(kernel_only, args.steps, 3)
    peak, peak_src = peaks()
    achieved = P * BYTES_PER_PAIR / (k_ms * 1e-3) / 1e9
    del out, ws, tar_flat, counts, pipe
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs, same run, same ranks (short step counts)
    extra = None
    if not args.no_extras:
        sys.path.insert(0, str(ROOT / "tools"))
        import bench_configs as bc
        extra = {}
        for name, fn in (("c3_train", lambda: bc.c3_train(pkg, dev, fused=True, dist=dist)),
                         ("c4_sharded", lambda: bc.c4_sharded(pkg, dev, dist=dist)),
                         ("c5_step", lambda: bc.c5_step(pkg, dev, dist=dist))):
            try:
                extra[name] = fn()
            except Exception as e:                      # an extra must never take the headline line down
                extra[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
            torch.cuda.empty_cache()

    if rank == 0:
        n = world
        value = n * P * args.steps / (ms_total * 1e-3)
        e2e_value = n * P * args.steps / (ms_e2e * 1e-3)
        copy_ms = ms_copy / args.steps
        ceil_gbs = n * h2d / (copy_ms * 1e-3) / 1e9
        e2e_gbs = n * h2d / (ms_e2e / args.steps * 1e-3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(n),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps,
                        "h2d_ceiling_gbs": ceil_gbs, "h2d_achieved_gbs": e2e_gbs, "frac_of_h2d_ceiling": e2e_gbs / ceil_gbs,
                        "h2d_ceiling_note": f"aggregate over {n} rank(s): the same pinned buffers ({h2d} B per rank per step) as "
                                            "plain cudaMemcpyAsync, all ranks at once, nothing else running, measured in "
                                            "this run (best of three timed blocks)",
                        "host_inputs": "pinned host tensors: fp32 v0 + binary cine masks as "
                                       + ("numpy.packbits bytes (1 bit per pixel" if args.mask_format == "bits"
                                          else "uint8 (1 B per pixel") +
                                       ", widened on the device, lossless); strain matrices back on the host, every "
                                       "step's result waited for inside the timed region (HostPipeline.submit / "
                                       "PipelineResult.get)",
                        "mask_format": args.mask_format},
                "gpu_launches": launches,
                "roofline": {"bound": "hbm", "kernel": "shoot_fwd_kernel<128,128,1024,clamp> (fused flat + 10 EPDiff steps + warp + strain)",
                             "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": measured_traffic(), "peak_source": peak_src, "kernel_ms": k_ms,
                             "kernel_ms_min_max": [k_min, k_max], "launches_timed": args.steps,
                             "algorithmic_bytes_per_launch": P * BYTES_PER_PAIR,
                             "note": "op-level algorithmic bytes (700*N per pair); the fused kernel keeps m/v on chip, "
                                     "so real DRAM traffic is far lower (see profiles/); traffic is null when the kernel "
                                     "sources changed since the ncu capture in profiles/traffic.json"}}
        if extra is not None:
            line["extra"] = extra
        if n == 1 and not args.no_cpu_baseline:
            if all_cpus:                                 # the CPU baseline gets every host core again
                os.sched_setaffinity(0, all_cpus)
            cpu_v, what, threads = cpu_pairs_per_s(pkg, args.cpu_sample_slices, 3)
            line["cpu_baseline"] = {"value": cpu_v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{args.cpu_sample_slices} slices x {T1} pairs of the same workload, "
                                              f"{what}, median of 3 passes"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
