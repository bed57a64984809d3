#!/usr/bin/env python
"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck)."""
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
dev = torch.device("cuda:0")
metric = pkg.FluidMetric((1.0, 0.1, 0.05))
for (B, T, H, W, S) in [(2, 3, 32, 32, 2), (1, 3, 64, 64, 2), (1, 2, 128, 128, 2), (1, 2, 64, 128, 2), (1, 2, 256, 256, 1)]:
    vol = pkg.synthetic.synthetic_masks(B, T, H, W).to(dev)
    v0 = pkg.synthetic.synthetic_v0(B * (T - 1), H, W, max_disp=3.0).to(dev).requires_grad_(True)
    sv, tv = pkg.data.split_vol_to_registration_pairs(vol, "Lagrangian", 3)
    out = pkg.shoot_warp_strain(v0, sv, tv, metric, num_steps=S)
    loss = out["deformed_source"].pow(2).mean() + out["strain_matrix"].pow(2).mean() + (out["velocity"] * out["momentum"]).sum()
    loss.backward()
    out = pkg.shoot_warp_strain(v0, sv, tv, metric, num_steps=S, loss_terms=True)      # loss epilogue + seedless adjoints
    (pkg.RegistrationReconstructionLoss(0.03, 0.1)(out, {"registration_target": tv}) + out["strain_matrix"].pow(2).mean()).backward()
    aug = pkg.augment.augment_batch(vol, out["strain_matrix"].detach(), None, 5, -3, 7)
    ids = [f"s{i // (T - 1)}" for i in range(B * (T - 1))]
    pkg.data.merge_data_of_same_slice_from_batch(
        {"slice_full_id": ids, "TOS": torch.zeros(len(ids), 126), "sector_LMA_labels": torch.zeros(len(ids), 126),
         "slice_LMA_label": torch.zeros(len(ids))}, {"displacement": out["displacement"].detach()}, 4, dev)
    u = out["displacement"].detach()
    I = torch.randn(B * (T - 1), 3, H, W, device=dev, requires_grad=True)
    ud = (3 * torch.randn_like(u)).requires_grad_(True)
    (pkg.interp(I, ud, 0.7, "zero").sum() + pkg.splat(I, ud, 0.7).sum()).backward()
    a = torch.randn_like(u, requires_grad=True)
    b = torch.randn_like(u, requires_grad=True)
    (pkg.Ad_star(a, b).sum() + pkg.jacobian_times_vectorfield(a, b, True, True).sum() + pkg.compose_disp_vel(a, b, -0.1).sum()).backward()
    m0 = metric.flat(u).requires_grad_(True)
    pkg.expmap(metric, m0, num_steps=S).sum().backward()
    pkg.sector_map(vol[:, 0, 0].contiguous())
    torch.cuda.synchronize()
    print("ok", B, T, H, W)
pipe = pkg.HostPipeline(3, 3, 32, 32, metric, num_steps=2, chunk_slices=2, device=dev)
pipe(pkg.synthetic.synthetic_v0(6, 32, 32).pin_memory(), pkg.synthetic.synthetic_masks(3, 3, 32, 32).pin_memory())
torch.cuda.synchronize()
print("sanitize_small done")
