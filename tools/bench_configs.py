#!/usr/bin/env python
"""Measurements of the other BASELINE.json configs (configs[2], [3], [4]).

`bench.py` imports these functions and appends their results to its JSON line as `extra.c3_train`,
`extra.c4_sharded` and `extra.c5_step` (same run, same ranks).  Stand-alone use:

  python tools/bench_configs.py c3      # configs[2]: C2 batch + EPDiff adjoint backward (training-mode gradients)
  python tools/bench_configs.py c3f     # same with the fused loss epilogue (per-pair loss terms from the kernel)
  python tools/bench_configs.py c4      # configs[3] shard: 16 slices x 50 frames 256x256 forward (cluster kernel)
  python tools/bench_configs.py c4o     # same through the op-level path (B2_FLAG_OPLEVEL)
  python tools/bench_configs.py c4t     # configs[3] shard, training step (forward with trajectory + fused adjoint)
  python tools/bench_configs.py c4s     # configs[3] full batch, 256 slices x 50 frames strong-sharded over the ranks
  python tools/bench_configs.py c5      # configs[4]: full training step with the stand-in nets
  (N GPUs: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_configs.py c4s|c5)

Prints one JSON line per config with pairs/s and the op-level roofline fraction.
"""
import contextlib
import json
import os
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

PARAMS = (1.0, 0.1, 0.05)
S_STEPS = 10


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    try:
        return float(json.loads(p.read_text())["hbm_gbs"])
    except Exception:
        return 6650.0


def bytes_fwd(H, W, S=S_STEPS):
    """Algorithmic bytes per frame-pair of the forward path: 4 (15 + 16 S) N."""
    return 4 * (15 + 16 * S) * H * W


def timed(fn, steps, warmup, barrier=None):
    for _ in range(warmup):
        fn()
    if barrier:
        barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def max_over_ranks(ms, dev, dist):
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def device_v0(pkg, P, H, W, seed, dev, max_disp=3.0, chunk=2048):
    """Smooth random velocities generated ON the device (data synthesis, outside every timed region):
    sharp(noise) rescaled to max |v0| = max_disp px per pair - the recipe of synthetic.synthetic_v0."""
    metric = pkg.FluidMetric(PARAMS)
    g = torch.Generator(device=dev).manual_seed(seed)
    out = torch.empty(P, 2, H, W, device=dev)
    with torch.no_grad():
        for i in range(0, P, chunk):
            n = min(chunk, P - i)
            v = metric.sharp(torch.randn(n, 2, H, W, device=dev, generator=g))
            mag = v.pow(2).sum(1).sqrt().amax(dim=(1, 2)).clamp(min=1e-12).view(n, 1, 1, 1)
            out[i:i + n] = v * (max_disp / mag)
    return out


def c3_train(pkg, dev, fused=True, steps=5, warmup=2, dist=None, B=64, T=25, H=128, W=128):
    """configs[2]: the C2 batch per GPU with the EPDiff adjoint: forward (trajectory saved) + gradient of the trainer
    loss w.r.t. v0 through strain, warp and shooting.  Contract bytes: 3 x forward (backward = 2 x forward)."""
    metric = pkg.FluidMetric(PARAMS)
    P = B * (T - 1)
    rank = dist.get_rank() if dist is not None else 0
    vol = pkg.synthetic.synthetic_masks(B, T, H, W, seed=2434 + 17 * rank).to(dev)
    v0 = device_v0(pkg, P, H, W, 5 + rank, dev)
    src_vol, tar_vol = pkg.data.split_vol_to_registration_pairs(vol, "Lagrangian", 3)
    Sgt = torch.zeros(B, 1, 126, 40, device=dev)
    vg = v0.clone().requires_grad_(True)
    crit = pkg.RegistrationReconstructionLoss(0.03, 0.1)

    def step():
        vg.grad = None
        out = pkg.shoot_warp_strain(vg, src_vol, tar_vol, metric, num_steps=S_STEPS, loss_terms=fused)
        loss = crit(out, {"registration_target": tar_vol}) + 1000.0 * torch.mean((out["strain_matrix"] - Sgt) ** 2)
        loss.backward()

    ms = max_over_ranks(timed(step, steps, warmup, dist.barrier if dist else None), dev, dist)
    world = dist.get_world_size() if dist is not None else 1
    nbytes = 3 * bytes_fwd(H, W)
    ach = P * nbytes / (ms * 1e-3) / 1e9
    return {"config": f"configs[2]: {B} slices x {T} frames {H}x{W} per GPU, forward + EPDiff adjoint backward"
                      + (" (fused loss epilogue)" if fused else ""),
            "pairs_per_gpu": P, "n_gpus": world, "ms_per_step": ms, "pairs_per_s": world * P / (ms * 1e-3),
            "roofline": {"achieved_GBps": ach, "peak_GBps": hbm_peak(), "frac": ach / hbm_peak(),
                         "algorithmic_bytes_per_pair": nbytes, "per": "GPU"}}


def c4_forward(pkg, dev, B=16, T=50, H=256, W=256, oplevel=False, train=False, steps=3, warmup=1):
    """configs[3] shard on one GPU: forward (cluster kernel or op-level path), or forward + adjoint."""
    metric = pkg.FluidMetric(PARAMS)
    P = B * (T - 1)
    vol = pkg.synthetic.synthetic_masks(B, T, H, W).to(dev)
    v0 = device_v0(pkg, P, H, W, 5, dev, chunk=784)
    src_vol, tar_vol = pkg.data.split_vol_to_registration_pairs(vol, "Lagrangian", 3)
    ctx = pkg.shooting.force_oplevel(fwd=True, bwd=True) if oplevel else contextlib.nullcontext()
    if train:
        Sgt = torch.zeros(B, 1, 126, 40, device=dev)
        vg = v0.clone().requires_grad_(True)
        crit = pkg.RegistrationReconstructionLoss(0.03, 0.1)

        def step():
            vg.grad = None
            out = pkg.shoot_warp_strain(vg, src_vol, tar_vol, metric, num_steps=S_STEPS, loss_terms=True)
            (crit(out, {"registration_target": tar_vol}) + 1000.0 * torch.mean((out["strain_matrix"] - Sgt) ** 2)).backward()
    else:
        def step():
            with torch.no_grad():
                pkg.shoot_warp_strain(v0, src_vol, tar_vol, metric, num_steps=S_STEPS)
    with ctx:
        ms = timed(step, steps, warmup)
    nbytes = (3 if train else 1) * bytes_fwd(H, W)
    ach = P * nbytes / (ms * 1e-3) / 1e9
    what = ("training step (forward + adjoint)" if train else "forward") + (", op-level path" if oplevel else ", cluster kernels")
    return {"config": f"configs[3] shard: {B} slices x {T} frames {H}x{W} {what}", "pairs": P, "ms_per_step": ms,
            "pairs_per_s": P / (ms * 1e-3),
            "roofline": {"achieved_GBps": ach, "peak_GBps": hbm_peak(), "frac": ach / hbm_peak(),
                         "algorithmic_bytes_per_pair": nbytes}}


def c4_sharded(pkg, dev, dist=None, B=256, T=50, H=256, W=256, steps=2, warmup=1):
    """configs[3]: the WHOLE batch (256 slices x 50 frames of 256x256 = 12 544 frame-pairs) strong-sharded by slice
    over the ranks (parallel.shard_slices): total work fixed, no data-path collective; time = max over ranks."""
    metric = pkg.FluidMetric(PARAMS)
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    b0, b1 = pkg.parallel.shard_slices(B, rank, world)
    nb, T1 = b1 - b0, T - 1
    vol = pkg.synthetic.synthetic_masks(nb, T, H, W, seed=2434 + b0, device=dev)    # this rank's slices, built on the GPU
    v0 = device_v0(pkg, nb * T1, H, W, 1000 + b0, dev, chunk=784)
    src_vol, tar_vol = pkg.data.split_vol_to_registration_pairs(vol, "Lagrangian", 3)
    keep = {}

    def step():
        keep["out"] = None        # free the previous outputs first: the caching allocator reuses the blocks (no cudaMalloc)
        with torch.no_grad():
            keep["out"] = pkg.shoot_warp_strain(v0, src_vol, tar_vol, metric, num_steps=S_STEPS)

    ms = max_over_ranks(timed(step, steps, warmup, dist.barrier if dist else None), dev, dist)
    assert all(torch.isfinite(v).all() for v in keep["out"].values())
    S_all = pkg.parallel.gather_strain_matrices(keep["out"]["strain_matrix"], B)     # inference gather (rank 0)
    P = B * T1
    ach = P * bytes_fwd(H, W) / (ms * 1e-3) / 1e9
    res = {"config": f"configs[3]: {B} slices x {T} frames {H}x{W} ({P} frame-pairs) sharded by slice over {world} GPU(s)",
           "scaling": "strong", "n_gpus": world, "pairs": P, "slices_per_gpu": nb, "ms_per_step": ms,
           "pairs_per_s": P / (ms * 1e-3),
           "roofline": {"achieved_GBps": ach, "peak_GBps": world * hbm_peak(), "frac": ach / (world * hbm_peak()),
                        "algorithmic_bytes_per_pair": bytes_fwd(H, W), "per": "job (N x peak)"},
           "peak_mem_GiB": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
    if rank == 0:
        res["gathered_strain_matrix_shape"] = list(S_all.shape)
    return res


def c5_step(pkg, dev, dist=None, steps=5, warmup=3, B=64, T=25, H=128, W=128, overlap=True):
    """configs[4]: per GPU a C2-shaped batch through the stand-in velocity net, the path (fused loss epilogue), the
    LMA net, the reference's three losses (weights 1 / 1000 / 0.005, configs/config.json:169,181,191), backward with
    the parameter-gradient all-reduce (NCCL) launched bucket by bucket from autograd hooks - the LMA bucket is in
    flight while the EPDiff adjoint kernel still runs - and an Adam step."""
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    # the stand-in networks are cuDNN library code outside the path: let them use the tensor cores
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.manual_seed(2434)                                      # same initial weights on every rank
    joint = pkg.build_model({"type": "JointRegisterStrainMatNet", "num_steps": S_STEPS, "fused_loss_terms": True}).to(dev)
    lma = pkg.build_model({"type": "NetStrainMat2LMA"}).to(dev)
    params = list(joint.parameters()) + list(lma.parameters())
    opt = torch.optim.Adam(params, lr=1e-4)
    vol = pkg.synthetic.synthetic_masks(B, T, H, W, seed=2434 + rank).to(dev)       # this rank's slices
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    Sgt = 0.05 * torch.randn(B, 1, 126, 40, device=dev, generator=g)
    tos = 60 * torch.rand(B, 126, device=dev, generator=g)
    crit = pkg.RegistrationReconstructionLoss(0.03, 0.1)
    # buckets in the order their gradients become ready in backward: LMA net first, velocity net last
    reducer = pkg.parallel.GradientAllReducer([list(lma.parameters()), list(joint.parameters())]) if overlap else None
    ncoll = [0]
    last = {}

    def step():
        opt.zero_grad(set_to_none=True)
        src_vol, tar_vol = pkg.data.split_vol_to_registration_pairs(vol, "Lagrangian", 3)
        out = joint.forward_volume(src_vol, tar_vol)
        pred_tos = lma(out["strain_matrix"])["TOS"]
        loss = 1.0 * crit(out, {"registration_target": tar_vol}) \
            + 1000.0 * torch.mean((out["strain_matrix"] - Sgt) ** 2) + 0.005 * torch.mean((pred_tos - tos) ** 2)
        loss.backward()
        ncoll[0] = reducer.finish() if reducer is not None else pkg.parallel.allreduce_gradients(params)
        opt.step()
        last["loss"] = loss

    ms = max_over_ranks(timed(step, steps, warmup, dist.barrier if dist else None), dev, dist)
    identical = True
    if world > 1:      # replicas must hold identical weights after identical averaged updates
        w = torch.cat([p.detach().reshape(-1) for p in params])
        ref = w.clone()
        dist.broadcast(ref, 0)
        identical = bool(torch.equal(w, ref))
    if reducer is not None:
        reducer.remove()
    P = B * (T - 1)
    return {"config": f"configs[4]: training step (velocity net + path + LMA net + 3 losses + gradient all-reduce + Adam), "
                      f"{B} slices x {T} frames {H}x{W} per GPU", "n_gpus": world, "ms_per_step": ms,
            "pairs_per_s": world * P / (ms * 1e-3), "collectives_per_step": ncoll[0],
            "allreduce": "bucketed, launched from autograd hooks (overlaps the adjoint)" if overlap else "after backward",
            "replicas_identical": identical, "loss": float(last["loss"].detach()), "tf32_nets": True}


def _init():
    import __graft_entry__ as g
    pkg = g.load_package()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    return pkg, dev, dist, rank


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "c3"
    pkg, dev, dist, rank = _init()
    if which in ("c3", "c3f"):
        res = c3_train(pkg, dev, fused=which == "c3f", dist=dist)
    elif which in ("c4", "c4o", "c4t"):
        res = c4_forward(pkg, dev, oplevel=which == "c4o", train=which == "c4t")
    elif which == "c4s":
        res = c4_sharded(pkg, dev, dist=dist)
    elif which == "c5":
        res = c5_step(pkg, dev, dist=dist)
    else:
        raise SystemExit(f"unknown config {which}")
    if rank == 0:
        print(json.dumps(res), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
