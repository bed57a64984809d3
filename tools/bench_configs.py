#!/usr/bin/env python
"""Secondary measurements for the other BASELINE.json configs (not the driver's bench line).

  python tools/bench_configs.py c3      # configs[2]: C2 batch + EPDiff adjoint backward (training-mode gradients)
  python tools/bench_configs.py c3f     # same with the fused loss epilogue (per-pair loss terms from the kernel)
  python tools/bench_configs.py c5      # configs[4]: full training step with the stand-in nets; N GPUs: torchrun --nproc-per-node N tools/bench_configs.py c5
  python tools/bench_configs.py c4      # configs[3]: 256x256, 50 frames (4-CTA cluster kernel; B2_NO_CLUSTER=1 = op-level path), a shard of 16 slices

Prints one JSON line per config with pairs/s and the op-level roofline fraction.
"""
import json
import os
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402

PARAMS = (1.0, 0.1, 0.05)


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def c5(pkg):
    """configs[4]: per GPU a C2-shaped batch through the stand-in velocity net, the path, the LMA net, the
    reference's three losses (weights 1 / 1000 / 0.005, configs/config.json:169,181,191), backward, ONE bucketed
    all-reduce of the parameter gradients (NCCL when launched under torchrun) and an Adam step."""
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, T, H, W, S = 64, 25, 128, 128, 10
    # the stand-in networks are cuDNN library code outside the path: let them use the tensor cores
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = os.environ.get("B2_C5_TF32", "1") == "1"
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32
    bf16 = os.environ.get("B2_C5_BF16", "0") == "1"
    torch.manual_seed(2434)                                      # same initial weights on every rank
    joint = pkg.build_model({"type": "JointRegisterStrainMatNet", "num_steps": S, "fused_loss_terms": True}).to(dev)
    lma = pkg.build_model({"type": "NetStrainMat2LMA"}).to(dev)
    params = list(joint.parameters()) + list(lma.parameters())
    opt = torch.optim.Adam(params, lr=1e-4)
    vol = pkg.synthetic.synthetic_masks(B, T, H, W, seed=2434 + rank).to(dev)       # this rank's slices
    Sgt = 0.05 * torch.randn(B, 1, 126, 40, device=dev)
    tos = 60 * torch.rand(B, 126, device=dev)
    crit = pkg.RegistrationReconstructionLoss(0.03, 0.1)
    ncoll = [0]

    def step():
        opt.zero_grad(set_to_none=True)
        src_vol, tar_vol = pkg.data.split_vol_to_registration_pairs(vol, "Lagrangian", 3)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16):
            out = joint.forward_volume(src_vol, tar_vol)
        pred_tos = lma(out["strain_matrix"])["TOS"]
        loss = 1.0 * crit(out, {"registration_target": tar_vol}) \
            + 1000.0 * torch.mean((out["strain_matrix"] - Sgt) ** 2) + 0.005 * torch.mean((pred_tos - tos) ** 2)
        loss.backward()
        ncoll[0] = pkg.parallel.allreduce_gradients(params)
        opt.step()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 10
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        # replicas must hold identical weights after identical averaged updates
        w = torch.cat([p.detach().reshape(-1) for p in params])
        ref = w.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(w, ref), "ranks diverged"
    if rank == 0:
        P = B * (T - 1)
        print(json.dumps({"config": f"configs[4]: training step (velocity net + path + LMA net + losses + all-reduce + Adam), "
                                    f"{B} slices x {T} frames 128x128 per GPU", "n_gpus": world, "ms_per_step": ms.item(),
                          "pairs_per_s": world * P / (ms.item() * 1e-3), "collectives_per_step": ncoll[0],
                          "loss": float(loss.detach()), "tf32_nets": torch.backends.cudnn.allow_tf32, "bf16_nets": bf16}))
    if world > 1:
        dist.destroy_process_group()


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "c3"
    pkg = g.load_package()
    if which == "c5":
        return c5(pkg)
    dev = torch.device("cuda:0")
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
    metric = pkg.FluidMetric(PARAMS)
    S = 10
    if which in ("c3", "c3f"):
        B, T, H, W = 64, 25, 128, 128
    else:
        B, T, H, W = 16, 50, 256, 256
    T1, P, N = T - 1, B * (T - 1), H * W
    vol = pkg.synthetic.synthetic_masks(B, T, H, W).to(dev)
    v0 = pkg.synthetic.synthetic_v0(P, H, W, seed=5, max_disp=3.0).to(dev)
    src_vol, tar_vol = pkg.data.split_vol_to_registration_pairs(vol, "Lagrangian", 3)
    bytes_fwd = 4 * (15 + 16 * S) * N
    if which in ("c3", "c3f"):
        Sgt = torch.zeros(B, 1, 126, 40, device=dev)
        vg = v0.clone().requires_grad_(True)
        fused = which == "c3f"            # loss epilogue: per-pair sums from the kernel, seedless adjoints
        crit = pkg.RegistrationReconstructionLoss(0.03, 0.1)

        def step():
            vg.grad = None
            out = pkg.shoot_warp_strain(vg, src_vol, tar_vol, metric, num_steps=S, loss_terms=fused)
            loss = crit(out, {"registration_target": tar_vol}) \
                + 1000.0 * torch.mean((out["strain_matrix"] - Sgt) ** 2)
            loss.backward()

        ms = timed(step, 5, 2)
        nbytes = 3 * bytes_fwd
        name = "configs[2]: C2 batch + EPDiff adjoint backward" + (" (fused loss epilogue)" if fused else "")
    else:
        def step():
            with torch.no_grad():
                pkg.shoot_warp_strain(v0, src_vol, tar_vol, metric, num_steps=S)

        ms = timed(step, 3, 1)
        nbytes = bytes_fwd
        name = f"configs[3] shard: {B} slices x {T} frames 256x256 forward ({'op-level path' if os.environ.get('B2_NO_CLUSTER') else 'cluster kernel'})"
    ach = P * nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"config": name, "pairs": P, "ms_per_step": ms, "pairs_per_s": P / (ms * 1e-3),
                      "roofline": {"achieved_GBps": ach, "peak_GBps": peak, "frac": ach / peak,
                                   "algorithmic_bytes_per_pair": nbytes}}))


if __name__ == "__main__":
    main()
