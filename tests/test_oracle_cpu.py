"""CPU tests: the oracle against the reference-generated golden vectors and the
invariants of SURVEY.md section 8c (the path itself is "parity unpinned")."""
import math

import numpy as np
import pytest
import torch

from conftest import relerr

F64 = torch.float64


def _rand(*s, seed=0, dtype=F64):
    return torch.randn(*s, generator=torch.Generator().manual_seed(seed), dtype=dtype)


# ----------------------------------------------------------------- golden vectors (reference code)
@pytest.mark.parametrize("method", ["Lagrangian", "Eulerian"])
@pytest.mark.parametrize("od", [2, 3])
def test_split_pairs_matches_reference(golden, oracle, pkg, method, od):
    vol = torch.from_numpy(golden["split_vol"])
    for impl in (oracle.split_vol_to_registration_pairs, pkg.data.split_vol_to_registration_pairs):
        s, t = impl(vol, method, od)
        assert np.array_equal(s.numpy(), golden[f"split_{method}_{od}_src"])
        assert np.array_equal(t.numpy(), golden[f"split_{method}_{od}_tar"])


def test_split_pairs_errors(oracle, pkg):
    for impl in (oracle.split_vol_to_registration_pairs, pkg.data.split_vol_to_registration_pairs):
        with pytest.raises(ValueError):
            impl(torch.zeros(1, 1, 3, 4, 4), "Bogus")
        with pytest.raises(AssertionError):
            impl(torch.zeros(1, 1, 1, 4, 4))


def test_align_frames_matches_reference(golden, oracle, pkg):
    a = golden["align_in"]
    for impl in (oracle.align_n_frames_to, pkg.data.align_n_frames_to):
        assert np.array_equal(impl(a, 3), golden["align_crop3"])
        assert np.array_equal(impl(a, 10), golden["align_pad10"])
        assert np.array_equal(impl(a, 9, frame_idx=1), golden["align_pad9_axis1"])
        assert np.array_equal(impl(a, 2, frame_idx=0), golden["align_crop2_axis0"])
    S = torch.from_numpy(a)
    assert np.array_equal(oracle.align_frames(S, 10).numpy(), golden["align_pad10"])
    assert np.array_equal(oracle.align_frames(S, 3).numpy(), golden["align_crop3"])


def test_fused_loss_class_matches_reference(golden, oracle, pkg):
    """The drop-in RegistrationReconstructionLoss (host logic of the loss epilogue): from tensors and from the
    per-pair terms it reproduces the value the reference's own LossCalculator gave for the golden prediction."""
    pred = {k[len("loss_pred_"):]: torch.from_numpy(golden[k]) for k in golden.files if k.startswith("loss_pred_")}
    target = {k[len("loss_target_"):]: torch.from_numpy(golden[k]) for k in golden.files if k.startswith("loss_target_")}
    want = float(golden["loss_part_registration_reconstruction"])
    crit = pkg.RegistrationReconstructionLoss(float(golden["loss_sigma"]), float(golden["loss_reg_weight"]))
    assert abs(crit(pred, target).item() - want) <= 1e-5 * abs(want)
    terms = oracle.path.registration_loss_terms(pred, target["registration_target"])
    assert terms.shape == (pred["velocity"].shape[0], 2)
    fused = crit({**pred, "registration_loss_terms": terms.float()}, target)
    assert abs(fused.item() - want) <= 1e-5 * abs(want)


def test_loss_boundary_matches_reference(golden, oracle):
    pred = {k[len("loss_pred_"):]: torch.from_numpy(golden[k]) for k in golden.files if k.startswith("loss_pred_")}
    target = {k[len("loss_target_"):]: torch.from_numpy(golden[k]) for k in golden.files if k.startswith("loss_target_")}
    rec = oracle.registration_reconstruction_loss(pred, target, float(golden["loss_sigma"]), float(golden["loss_reg_weight"]))
    assert abs(rec.item() - float(golden["loss_part_registration_reconstruction"])) <= 1e-5 * abs(rec.item())
    sup = torch.mean((pred["strainmat"] - target["strainmat"]) ** 2)
    tos = torch.mean((pred["TOS"] - target["TOS"]) ** 2)
    w = golden["loss_weights"]
    total = w[0] * rec + w[1] * sup + w[2] * tos
    assert abs(total.item() - float(golden["loss_total"])) <= 1e-5 * abs(total.item())


def test_reference_constants(golden, oracle):
    assert int(golden["n_sectors"]) == oracle.N_SECTORS == 126
    assert int(golden["n_strain_frames"]) == 40
    assert int(golden["seed"]) == 2434


# ----------------------------------------------------------------- operator invariants (8c)
@pytest.mark.parametrize("bg", ["clamp", "zero"])
def test_interp_identity_and_shift(oracle, bg):
    conv = oracle.Conventions(background=bg)
    I = _rand(2, 3, 9, 7)
    z = torch.zeros(2, 2, 9, 7, dtype=F64)
    assert torch.equal(oracle.interp(I, z, conv=conv), I)
    u = torch.zeros(2, 2, 9, 7, dtype=F64)
    u[:, 0] = 1.0
    u[:, 1] = -2.0
    out = oracle.interp(I, u, conv=conv)
    rr = torch.arange(9).view(9, 1) + 1
    cc = torch.arange(7).view(1, 7) - 2
    if bg == "clamp":
        ref = I[:, :, rr.clamp(0, 8), cc.clamp(0, 6)]
    else:
        ok = ((rr >= 0) & (rr < 9) & (cc >= 0) & (cc < 7)).to(F64)
        ref = I[:, :, rr.clamp(0, 8), cc.clamp(0, 6)] * ok
    assert torch.allclose(out, ref, atol=1e-14)


@pytest.mark.parametrize("bg", ["clamp", "zero"])
def test_interp_splat_adjoint(oracle, bg):
    conv = oracle.Conventions(background=bg)
    I, J, u = _rand(3, 2, 10, 8, seed=1), _rand(3, 2, 10, 8, seed=2), 3 * _rand(3, 2, 10, 8, seed=3)
    lhs = (oracle.interp(I, u, 0.7, conv) * J).sum()
    rhs = (I * oracle.splat(J, u, 0.7, conv=conv)).sum()
    assert abs(lhs - rhs) < 1e-10
    out, w = oracle.splat(J, u, 0.7, need_weights=True, conv=conv)
    assert torch.allclose(w, oracle.splat(torch.ones(3, 1, 10, 8, dtype=F64), u, 0.7, conv=conv))


def test_interp_batch_broadcast(oracle):
    I, u = _rand(1, 2, 6, 6, seed=4), _rand(3, 2, 6, 6, seed=5)
    out = oracle.interp(I, u)
    assert out.shape == (3, 2, 6, 6)
    assert torch.allclose(out[1:2], oracle.interp(I, u[1:2]))
    with pytest.raises(ValueError):
        oracle.interp(_rand(2, 1, 6, 6), u)


def test_fluid_metric_invariants(oracle):
    m = oracle.FluidMetric((1.0, 0.1, 0.05))
    v, w = _rand(2, 2, 16, 12, seed=6), _rand(2, 2, 16, 12, seed=7)
    assert relerr(m.sharp(m.flat(v)), v) < 1e-12
    assert abs((m.flat(v) * w).sum() - (v * m.flat(w)).sum()) < 1e-9
    const = torch.ones(1, 2, 16, 12, dtype=F64) * torch.tensor([2.0, -3.0], dtype=F64).view(1, 2, 1, 1)
    assert relerr(m.flat(const), 0.05 * const) < 1e-12
    # Fourier form == periodic finite-difference stencil
    a, b, g = 1.0, 0.1, 0.05
    lap = lambda f: sum(torch.roll(f, s, d) for s in (1, -1) for d in (2, 3)) - 4 * f
    d = lambda f, dim: 0.5 * (torch.roll(f, -1, dim) - torch.roll(f, 1, dim))
    dd = lambda f, dim: torch.roll(f, -1, dim) - 2 * f + torch.roll(f, 1, dim)
    v0, v1 = v[:, :1], v[:, 1:]
    L0 = g * v0 - a * lap(v0) - b * (dd(v0, 2) + d(d(v1, 3), 2))
    L1 = g * v1 - a * lap(v1) - b * (d(d(v0, 2), 3) + dd(v1, 3))
    assert relerr(m.flat(v), torch.cat([L0, L1], 1)) < 1e-12
    with pytest.raises(ValueError):
        oracle.FluidMetric((1.0, 0.1, 0.0))


def test_expmap_properties(oracle):
    m = oracle.FluidMetric((1.0, 0.1, 0.05))
    z = torch.zeros(1, 2, 16, 16, dtype=F64)
    assert torch.equal(oracle.expmap(m, z, num_steps=4), z)
    v0 = m.sharp(_rand(1, 2, 32, 32, seed=8))
    v0 = 1e-3 * v0 / v0.abs().max()
    m0 = m.flat(v0)
    u = oracle.expmap(m, m0, num_steps=5)
    assert relerr(u, -m.sharp(m0)) < 1e-2          # first order: u ~ -T*sharp(m0)
    # <m_t, v_t> approximately conserved along the geodesic (smooth momentum; bilinear
    # resampling of a rough momentum is dissipative).  The det(I+Du) variant (D3 alt) is the
    # exact coadjoint action of a density and conserves better than the default.
    vs = m.sharp(m.sharp(v0))
    vs = 0.3 * vs / vs.abs().max()
    drift = {}
    for det in (False, True):
        u, traj = oracle.expmap(m, m.flat(vs), num_steps=40, trajectory=True,
                                conv=oracle.Conventions(adstar_det=det))
        e = [float((mm * vv).sum()) for _, mm, vv in traj]
        drift[det] = max(e) / min(e)
    assert drift[False] < 1.03 and drift[True] < 1.01


@pytest.mark.parametrize("fn", ["interp", "splat", "jtv", "jtvT", "adstar", "compose", "flat", "sharp", "strain"])
def test_oracle_gradcheck(oracle, fn):
    torch.manual_seed(0)
    H = W = 8
    a = _rand(2, 2, H, W, seed=11).requires_grad_(True)
    b = (0.8 * _rand(2, 2, H, W, seed=12)).requires_grad_(True)
    m = oracle.FluidMetric((1.0, 0.1, 0.05))
    if fn == "interp":
        f = lambda I, u: oracle.interp(I, u, 0.9)
    elif fn == "splat":
        f = lambda I, u: oracle.splat(I, u, 0.9)
    elif fn == "jtv":
        f = lambda v, w: oracle.jacobian_times_vectorfield(v, w, True, False)
    elif fn == "jtvT":
        f = lambda v, w: oracle.jacobian_times_vectorfield(v, w, False, True)
    elif fn == "adstar":
        f = lambda mm, u: oracle.Ad_star(u, mm)
    elif fn == "compose":
        f = lambda u, v: oracle.compose_disp_vel(u, v, -0.3)
    elif fn == "flat":
        f = lambda v, _: m.flat(v)
    elif fn == "sharp":
        f = lambda v, _: m.sharp(v)
    else:
        c0 = torch.tensor([3.3, 3.9], dtype=F64)
        c1 = torch.tensor([3.6, 4.2], dtype=F64)
        f = lambda _, u: oracle.strain_ecc(0.3 * u, c0, c1)[0]
    assert torch.autograd.gradcheck(f, (a, b), eps=1e-6, atol=1e-5, rtol=1e-4, nondet_tol=1e-12)


# ----------------------------------------------------------------- strain + sectors
def test_sector_classifier_exact(oracle):
    n = 126
    for k in range(n):
        ang = 2 * math.pi * (k + 0.5) / n
        dr, dc = round(1e6 * math.sin(ang)), round(1e6 * math.cos(ang))
        assert oracle.strain.classify_directions(np.array([dr]), np.array([dc]), n)[0] == k
    tab = oracle.sector_boundaries(n)
    # a direction exactly on boundary k belongs to sector k (closed at the lower edge)
    got = oracle.strain.classify_directions(tab[:, 0] * 7, tab[:, 1] * 7, n)
    assert np.array_equal(got, np.arange(n))
    assert oracle.strain.classify_directions(np.array([0]), np.array([0]), n)[0] == -1
    # table entries are not near rounding ties (so C llrint and numpy rint agree on any libm)
    k = np.arange(n)
    fr = np.abs(np.stack([np.sin(2 * np.pi * k / n), np.cos(2 * np.pi * k / n)]) * (1 << 20))
    assert (np.abs(fr - np.floor(fr) - 0.5) > 1e-6).all()


def test_sector_rotation_convention(oracle):
    """Rotating the content by +360/126 deg in theta = atan2(drow, dcol) raises the sector id by one
    (/root/reference/modules/data/augmentation/affine.py:56-78: rotate by -n*360/126 <-> roll rows by +n)."""
    n = 126
    rng = np.random.default_rng(0)
    ang = rng.uniform(0, 2 * np.pi, 500)
    rad = rng.uniform(1e5, 1e6, 500)
    d = lambda a: (np.rint(rad * np.sin(a)).astype(np.int64), np.rint(rad * np.cos(a)).astype(np.int64))
    k0 = oracle.strain.classify_directions(*d(ang), n)
    k1 = oracle.strain.classify_directions(*d(ang + 5 * 2 * np.pi / n), n)
    frac = (ang / (2 * np.pi / n)) % 1.0
    safe = (frac > 1e-3) & (frac < 1 - 1e-3)
    assert np.array_equal(k1[safe], (k0[safe] + 5) % n)


def _disc_masks(B, T1, H, W, r_in=10.0, r_out=29.0):
    rr = torch.arange(H, dtype=torch.float32).view(H, 1) - (H - 1) / 2
    cc = torch.arange(W, dtype=torch.float32).view(1, W) - (W - 1) / 2
    rad = torch.sqrt(rr * rr + cc * cc)
    m = ((rad >= r_in) & (rad <= r_out)).float()
    return m.expand(B, H, W).contiguous(), m.expand(B, T1, H, W).contiguous()


def test_strain_rigid_and_scaling(oracle):
    B, T1, H, W = 1, 2, 64, 64
    mask0, tar = _disc_masks(B, T1, H, W)
    rr = torch.arange(H, dtype=F64).view(H, 1).expand(H, W) - (H - 1) / 2
    cc = torch.arange(W, dtype=F64).view(1, W).expand(H, W) - (W - 1) / 2
    # rigid rotation + translation of the inverse map: X = R x + t  ->  zero strain
    th = 0.2
    u_rot = torch.stack([math.cos(th) * rr - math.sin(th) * cc - rr + 1.5,
                         math.sin(th) * rr + math.cos(th) * cc - cc - 0.5])
    # uniform radial scaling about the centroid: x = s X  ->  F = s I, Ecc = (s^2-1)/2
    s = 1.1
    u_sc = torch.stack([rr / s - rr, cc / s - cc])
    u = torch.stack([u_rot, u_sc]).unsqueeze(0)               # (1,2,2,H,W)
    S, cnt = oracle.strain_matrix(u, tar, mask0, n_frames=None, return_counts=True)
    assert S.shape == (1, 1, 126, 2)
    assert (cnt > 0).all()
    assert S[0, 0, :, 0].abs().max() < 1e-12
    assert (S[0, 0, :, 1] - (s * s - 1) / 2).abs().max() < 1e-12
    # frame alignment to 40 = edge padding of the last column
    S40 = oracle.strain_matrix(u, tar, mask0, n_frames=40)
    assert S40.shape == (1, 1, 126, 40)
    assert torch.equal(S40[..., 5], S40[..., 1])
    # empty mask -> zeros
    Z = oracle.strain_matrix(u, torch.zeros_like(tar), mask0, n_frames=None)
    assert Z.abs().max() == 0


def test_forward_volume_contract(oracle):
    """Keys/shapes the trainer reads (joint_registration_strainmat_LMA.py:314-318) and the loss accepts."""
    B, T, H, W = 2, 4, 32, 32
    g = torch.Generator().manual_seed(3)
    vol = (torch.rand(B, 1, T, H, W, generator=g) > 0.5).float()
    src_vol, tar_vol = oracle.split_vol_to_registration_pairs(vol, "Lagrangian", 3)
    v0 = 0.5 * torch.randn(B * (T - 1), 2, H, W, generator=g)
    m = oracle.FluidMetric((1.0, 0.1, 0.05))
    out = oracle.forward_volume(m.sharp(v0), src_vol, tar_vol, m, num_steps=3)
    assert out["strain_matrix"].shape == (B, 1, 126, 40)
    assert out["deformed_source"].shape == tar_vol.shape
    assert out["velocity"].shape == out["momentum"].shape == (B * (T - 1), 2, H, W)
    loss = oracle.registration_reconstruction_loss(out, {"registration_target": tar_vol})
    assert torch.isfinite(loss)


@pytest.mark.parametrize("cfg", [(2, 3, 32, 32, 3), (1, 4, 64, 64, 6), (1, 3, 32, 64, 2)])
def test_c_oracle_matches_torch(oracle, pkg, cfg):
    """The two independently written oracles (plain C with its own FFT / torch) agree on every output."""
    from oracle import c_oracle
    B, T, H, W, S = cfg
    vol = pkg.synthetic.synthetic_masks(B, T, H, W)
    v0 = pkg.synthetic.synthetic_v0(B * (T - 1), H, W, seed=41, max_disp=3.0)
    sv, tv = oracle.split_vol_to_registration_pairs(vol, "Lagrangian", 3)
    ref = oracle.forward_volume(v0, sv, tv, oracle.FluidMetric((1.0, 0.1, 0.05)), S)
    out = c_oracle.forward_volume(v0, vol, (1.0, 0.1, 0.05), S, nthreads=2)
    for k in ("momentum", "velocity", "displacement", "deformed_source", "strain_matrix"):
        assert out[k].shape == ref[k].shape
        assert relerr(out[k], ref[k]) < 1e-5, f"{k}: {relerr(out[k], ref[k]):.2e}"
    import ctypes
    tab = (ctypes.c_int32 * 252)()
    c_oracle.lib().b2o_sector_table(126, tab)
    assert np.array_equal(np.array(list(tab)).reshape(126, 2), oracle.sector_boundaries(126))


def test_oracles_constant_velocity_is_a_translation(oracle, pkg):
    """Known answer that pins both oracles to the mathematics instead of to each other: a constant initial velocity
    c is a fixed point of the EPDiff flow, so u = -T c, velocity = c, momentum = gamma c, the warped source is the
    source shifted by the integer c (edge clamp) and the strain of a translation vanishes."""
    from oracle import c_oracle
    B, T, H, W, S = 2, 3, 32, 32, 10
    par = (1.0, 0.1, 0.05)
    vol = pkg.synthetic.synthetic_masks(B, T, H, W)
    sv, tv = oracle.split_vol_to_registration_pairs(vol, "Lagrangian", 3)
    shifts = torch.tensor([[2.0, -3.0], [0.0, 1.0], [-1.0, -2.0], [3.0, 0.0]])
    v0 = shifts.view(4, 2, 1, 1).expand(4, 2, H, W).contiguous()
    r, c = torch.arange(H), torch.arange(W)
    for out in (oracle.forward_volume(v0, sv, tv, oracle.FluidMetric(par), S),
                c_oracle.forward_volume(v0, vol, par, S, nthreads=2)):
        assert (out["displacement"] + v0).abs().max() < 1e-5
        assert (out["velocity"] - v0).abs().max() < 1e-5
        assert (out["momentum"] - par[2] * v0).abs().max() < 1e-6
        for p in range(4):
            a, b = int(shifts[p, 0]), int(shifts[p, 1])
            want = vol[p // (T - 1), 0, 0][(r - a).clamp(0, H - 1)][:, (c - b).clamp(0, W - 1)]
            assert (out["deformed_source"][p // (T - 1), 0, p % (T - 1)] - want).abs().max() < 1e-4
        assert out["strain_matrix"].abs().max() < 1e-5


# ----------------------------------------------------------------- augmentation (reference affine.py)
@pytest.fixture(scope="module")
def aug_golden():
    import pathlib
    return np.load(pathlib.Path(__file__).resolve().parent / "golden" / "ref_augment.npz")


def test_augment_translate_matches_reference(aug_golden, oracle):
    """np.roll translation of masks (H,W,T layout in the reference, (B,1,T,H,W) here); strain / TOS untouched."""
    g = aug_golden
    mask = g["mask"]                                           # (H,W,T)
    vol = np.moveaxis(mask, -1, 0)[None, None]                 # (1,1,T,H,W)
    for i, (ty, tx) in enumerate(g["shifts"]):
        got = oracle.augment.rotate_translate_volume(vol, [0], [ty], [tx])
        assert np.array_equal(np.moveaxis(got[0, 0], 0, -1), g[f"translate_{i}_mask"])
        assert np.array_equal(g[f"translate_{i}_strain"], g["strain"]) and np.array_equal(g[f"translate_{i}_tos"], g["tos"])


def test_augment_rotate_convention_matches_reference(aug_golden, oracle):
    """Angle handed to skimage and the strain/TOS roll of affine.py:56,74,78; order=0 nearest neighbour."""
    g = aug_golden
    assert int(g["skrotate_order"]) == 0
    for i, n in enumerate(g["rot_n"]):
        assert oracle.augment.rotation_angle_degree(int(n)) == float(g["rot_angle_degree"][i])
        assert np.array_equal(oracle.augment.roll_rows(g["strain"][None, None], [n])[0, 0], g[f"rotate_{i}_strain"])
        assert np.array_equal(oracle.augment.roll_rows(g["tos"][None], [n])[0], g[f"rotate_{i}_tos"])


def test_augment_rotation_properties(oracle, pkg):
    """The restated skimage rotation: identity at n = 0 and n = 126, quarter turns are exact index permutations,
    and rotating a mask by n sectors shifts the (integer) sector ids of its pixels by n - the equivariance
    affine.py:56-78 relies on when it rolls the strain rows."""
    rng = np.random.default_rng(7)
    img = rng.random((3, 16, 16)).astype(np.float32)
    A = oracle.augment
    for n in (0, 126, -126):
        assert np.array_equal(A.rotate_nearest(img, A.rotate_matrix(A.rotation_angle_degree(n), 16, 16)), img)
    # skimage's angle is counter-clockwise on the displayed image: +90 deg == np.rot90(k=1) for a square image
    assert np.array_equal(A.rotate_nearest(img, A.rotate_matrix(90.0, 16, 16)), np.rot90(img, 1, axes=(-2, -1)))
    assert np.array_equal(A.rotate_nearest(img, A.rotate_matrix(-90.0, 16, 16)), np.rot90(img, -1, axes=(-2, -1)))
    # sector equivariance on an annulus, centroid at the rotation centre
    H = W = 128
    r, c = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    rad = np.hypot(r - 63.5, c - 63.5)
    ring = ((rad > 20) & (rad < 40)).astype(np.float32)
    ids0 = oracle.sector_map(torch.from_numpy(ring)[None], 126)[0].numpy()
    for n in (1, 5, 31):
        m = A.rotate_matrix(A.rotation_angle_degree(n), H, W)
        moved = A.rotate_nearest(ids0.astype(np.float32) + 1.0, m) - 1.0      # carry the ids with the pixels
        sel = (moved >= 0) & (ring > 0) & (A.rotate_nearest(ring, m) > 0)
        ids_new = oracle.sector_map(torch.from_numpy(A.rotate_nearest(ring, m))[None], 126)[0].numpy()
        d = (ids_new[sel] - moved[sel]) % 126
        # nearest-neighbour resampling moves a pixel by up to half a pixel: ids within one sector of the exact shift
        assert np.all((d == n % 126) | (d == (n - 1) % 126) | (d == (n + 1) % 126))
        assert np.mean(d == n % 126) > 0.8


# ----------------------------------------------------------------- slice regrouping (reference trainer helper)
def _regroup_inputs(g):
    batch = {"slice_full_id": [str(x) for x in g["ids"]], "TOS": torch.from_numpy(g["TOS"]),
             "sector_LMA_labels": torch.from_numpy(g["sector_LMA_labels"]),
             "slice_LMA_label": torch.from_numpy(g["slice_LMA_label"])}
    return batch, {"displacement": torch.from_numpy(g["displacement"])}


def check_regroup_against_golden(g, merge, F):
    """Compare per slice id: the reference iterates a set, so its slice order is arbitrary."""
    batch, pred = _regroup_inputs(g)
    r = merge(batch, pred, F)
    ref_ids = [str(x) for x in g[f"F{F}_ids"]]
    assert sorted(r["batch_slice_full_ids"]) == sorted(ref_ids)
    for j, sid in enumerate(r["batch_slice_full_ids"]):
        k = ref_ids.index(sid)
        assert np.array_equal(r["pred_displacement_fields"][j].cpu().numpy(), g[f"F{F}_fields"][k]), (F, sid)
        assert np.array_equal(r["TOS"][j].cpu().numpy(), g[f"F{F}_TOS"][k])
        assert np.array_equal(r["sector_LMA_labels"][j].cpu().numpy(), g[f"F{F}_labels"][k])
        assert int(r["slice_LMA_label"][j]) == int(g[f"F{F}_slice_label"][k])


@pytest.mark.parametrize("F", [3, 4, 7])
def test_regroup_matches_reference(oracle, F):
    import pathlib
    g = np.load(pathlib.Path(__file__).resolve().parent / "golden" / "ref_regroup.npz")
    check_regroup_against_golden(g, oracle.path.merge_data_of_same_slice_from_batch, F)


def test_sector_frame_matches_reference_mesh(oracle):
    """theta0 + direction of the sector classifier against the reference's own 126-sector mesh: the centres of the
    mid-wall faces of ``spl2patchSA`` (DENSE_utils.py:177-295; golden made by importing the reference function) must
    fall into sector k = their mesh index, for every start angle and both numbering directions."""
    import pathlib
    g = np.load(pathlib.Path(__file__).resolve().parent / "golden" / "ref_sectors.npz")
    o = g["origin_xy"]
    assert int(g["n_cases"]) == 6
    for i in range(int(g["n_cases"])):
        c = g[f"case{i}_midwall_centers_xy"]                       # (126, 2) as (x, y) = (col, row)
        theta0, cw = float(g[f"case{i}_theta0"]), bool(g[f"case{i}_clockwise"])
        dr = np.rint(1000 * (c[:, 1] - o[1])).astype(np.int64)
        dc = np.rint(1000 * (c[:, 0] - o[0])).astype(np.int64)
        k = oracle.strain.classify_directions(dr, dc, 126, theta0, cw)
        assert np.array_equal(k, np.arange(126)), (i, theta0, cw)
        # 18 segments x 7 samples (DENSE_utils.py:178-188): AHA-style segment id of sector k is k // 7 + 1
        assert np.array_equal(g[f"case{i}_sectorid"], np.arange(126) // 7 + 1)
    # the default frame is the case theta0 = 0, clockwise: bit-identical to the frameless call
    dr, dc = np.meshgrid(np.arange(-40, 41), np.arange(-40, 41), indexing="ij")
    assert np.array_equal(oracle.strain.classify_directions(dr, dc, 126),
                          oracle.strain.classify_directions(dr, dc, 126, 0.0, True))


def test_sector_frame_equivariance(oracle):
    """Shifting theta0 by j sector widths rolls the sector ids by -j; flipping the direction mirrors them."""
    dr, dc = np.meshgrid(np.arange(-30, 31), np.arange(-30, 31), indexing="ij")
    base = oracle.strain.classify_directions(dr, dc, 126, 0.4, True)
    ok = base >= 0
    w = 2 * np.pi / 126
    for j in (1, 5, 125):
        sh = oracle.strain.classify_directions(dr, dc, 126, 0.4 + j * w, True)
        # boundaries are rounded to Q20 independently per table: pixels sitting exactly on a boundary may differ
        frac = np.mean(((base - j) % 126)[ok] == sh[ok])
        assert frac > 0.999, (j, frac)
    ccw = oracle.strain.classify_directions(dr, dc, 126, 0.4, False)
    assert np.array_equal(ccw[ok], 125 - base[ok])


def test_c_oracle_sector_frame_matches_torch(oracle):
    """Both oracles agree on the strain matrix in a rotated / flipped sector frame."""
    from oracle import c_oracle
    import __graft_entry__ as g
    pkg = g.load_package()
    B, T, H, W, S = 3, 3, 32, 32, 2
    vol = pkg.synthetic.synthetic_masks(B, T, H, W)
    v0 = pkg.synthetic.synthetic_v0(B * (T - 1), H, W, seed=5, max_disp=2.0)
    th, cw = [0.7, -2.0, 0.0], [True, False, False]
    src_vol, tar_vol = oracle.split_vol_to_registration_pairs(vol, "Lagrangian", 3)
    a = oracle.forward_volume(v0, src_vol, tar_vol, oracle.FluidMetric((1.0, 0.1, 0.05)), S, theta0=th, clockwise=cw)
    b = c_oracle.forward_volume(v0, vol, (1.0, 0.1, 0.05), S, theta0=th, clockwise=cw)
    d = oracle.forward_volume(v0, src_vol, tar_vol, oracle.FluidMetric((1.0, 0.1, 0.05)), S)
    assert (a["strain_matrix"] - b["strain_matrix"]).abs().max() < 1e-5 * a["strain_matrix"].abs().max()
    assert (a["strain_matrix"] - d["strain_matrix"]).abs().max() > 1e-3 * a["strain_matrix"].abs().max()
    # slice 2 uses theta0 = 0 counter-clockwise: rows are the default rows reversed
    assert torch.equal(a["strain_matrix"][2].flip(1), d["strain_matrix"][2])
