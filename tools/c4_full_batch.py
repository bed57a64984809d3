import sys, json, torch, time
sys.path.insert(0, '/root/repo')
import __graft_entry__ as g
pkg = g.load_package()
dev = torch.device('cuda:0')
B, T, H, W, S = 256, 50, 256, 256, 10     # BASELINE configs[3], the whole batch on one GPU
P = B * (T - 1)
t0 = time.time()
vol = pkg.synthetic.synthetic_masks(B, T, H, W).to(dev)
# smooth velocity field generated on the GPU in chunks (data synthesis, not the measured path)
v0 = torch.empty(P, 2, H, W, device=dev)
for i in range(0, P, 784):
    v0[i:i + 784] = pkg.synthetic.synthetic_v0(min(784, P - i), H, W, seed=5 + i, max_disp=3.0, device='cpu').to(dev)
print('inputs ready', time.time() - t0, 's; v0 GiB', v0.numel() * 4 / 2**30)
sv, tv = pkg.data.split_vol_to_registration_pairs(vol, 'Lagrangian', 3)
metric = pkg.FluidMetric((1.0, 0.1, 0.05))
torch.cuda.synchronize()
with torch.no_grad():
    # warm-up: the first call cudaMallocs 25 GB of outputs inside the caching allocator; time only calls that reuse them
    out = pkg.shoot_warp_strain(v0, sv, tv, metric, num_steps=S)
    torch.cuda.synchronize()
    for it in range(3):
        out = None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = pkg.shoot_warp_strain(v0, sv, tv, metric, num_steps=S)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(json.dumps({'config': 'configs[3] full batch on 1 GPU', 'pairs': P, 'ms': ms, 'pairs_per_s': P / ms * 1e3,
                          'peak_mem_GiB': torch.cuda.max_memory_allocated() / 2**30}))
assert all(torch.isfinite(v).all() for v in out.values())
# shard independence at full size: slice block recomputed alone reproduces the same rows
sub = pkg.shoot_warp_strain(v0[:2 * (T - 1)], sv[:2], tv[:2], metric, num_steps=S)
print('shard identical:', torch.equal(sub['displacement'], out['displacement'][:2 * (T - 1)]), torch.equal(sub['strain_matrix'], out['strain_matrix'][:2]))
