"""Parity of the CUDA path with the fixtures recorded from the reference gicp.py (2-D, fp64) and
with the CPU oracle (3-D).  Every call goes through the C ABI (generalized_icp_b200.engine ->
ctypes -> libgicp_b200.so).  Needs a B200: run with -m gpu."""
import numpy as np
import pytest

from conftest import golden_names

pytestmark = pytest.mark.gpu

ALL = golden_names()
SMALL = [n for n in ALL if n != "config1_seed4"]


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _engine2d(g, torch, storage="f64"):
    from generalized_icp_b200.engine import GicpEngine
    eng = GicpEngine(2, storage)
    eng.set_params(k=6, max_iterations=int(g["max_iterations"]), tolerance=float(g["tolerance"]),
                   max_distance_correspondence=float(g["d_max"]), max_distance_nearest_neighbors=float(g["r_knn"]))
    dt = torch.float64 if storage == "f64" else torch.float32
    eng.set_target(torch.as_tensor(g["tgt"], dtype=dt, device="cuda"))
    eng.set_source(torch.as_tensor(g["src"], dtype=dt, device="cuda"))
    return eng


@pytest.mark.parametrize("name", ALL)
def test_knn_indices_bit_exact(name, golden, torch_cuda):
    """gicp.py:24-25: neighbour lists, same order, same tie-break (fixtures are tie-free)."""
    g = golden(name)
    eng = _engine2d(g, torch_cuda)
    for which, key, n in ((0, "src_knn", len(g["src"])), (1, "tgt_knn", len(g["tgt"]))):
        idx, dist = eng.knn(which)
        want = np.where(g[key] == n, -1, g[key])
        assert np.array_equal(idx.cpu().numpy(), want)
        d = dist.cpu().numpy()
        assert np.all(np.isinf(d[want < 0]))
        assert np.all(np.diff(np.where(np.isinf(d), 1e300, d), axis=1) >= 0)


@pytest.mark.parametrize("name", ALL)
def test_covariances(name, golden, torch_cuda):
    """gicp.py:5-35 (target once, source at T=I) and gicp.py:120 (per-iteration source covariances)."""
    g = golden(name)
    eng = _engine2d(g, torch_cuda)
    assert np.abs(eng.covariances(1).cpu().numpy() - g["tgt_cov"]).max() < 1e-9
    assert np.abs(eng.covariances(0).cpu().numpy() - g["src_cov0"]).max() < 1e-9
    n = len(g["all_src_cov"])
    Ts = g["all_T"][:n].reshape(n, 1, 3, 3)
    got = eng.source_covariances_at(Ts).cpu().numpy()
    assert np.abs(got - g["all_src_cov"]).max() < 1e-9


@pytest.mark.parametrize("name", SMALL)
def test_correspondences_weights_loss_teacher_forced(name, golden, torch_cuda):
    """Feed the reference's own T_k: 1-NN indices bit-exact (gicp.py:132-138), W (gicp.py:145), and the
    value (gicp.py:52-58) and gradient (gicp.py:60-76) of the frozen inner objective at the reference's
    x0 / xopt, both from the reduced form K3 accumulates."""
    from generalized_icp_b200.engine import reduced_form_grad2d, reduced_form_loss
    g = golden(name)
    eng = _engine2d(g, torch_cuda)
    for k in range(len(g["it_fopt"])):
        T = g["all_T"][k]
        idx, dist, W = eng.correspond(T)
        idx = idx.cpu().numpy()
        assert np.array_equal(idx, g["nn_idx"][k])
        m = idx >= 0
        assert np.abs(dist.cpu().numpy()[m] - g["nn_dist"][k][m]).max() < 1e-10
        assert np.abs(W.cpu().numpy() - g["it_W"][k]).max() < 1e-12
        red = eng.normal_equations(T)[0]
        assert int(round(red[25])) == int(m.sum())
        for xk, lk in (("it_x0", "it_loss_at_x0"), ("it_xopt", "it_loss_at_xopt")):
            x = g[xk][k]
            Te = np.eye(3)
            Te[:2, :2] = [[np.cos(x[2]), -np.sin(x[2])], [np.sin(x[2]), np.cos(x[2])]]
            Te[:2, 2] = x[:2]
            want = float(g[lk][k])
            got = reduced_form_loss(red, 2, T, Te)
            assert abs(got - want) <= 1e-9 * max(1.0, abs(want)), (k, xk, got, want)
            # grad_loss (gicp.py:60-76) at the same point, from the same 32 doubles
            gw = g[lk.replace("loss", "grad")][k]
            gg = reduced_form_grad2d(red, T, x)
            assert np.abs(gg - gw).max() <= 1e-7 * max(1.0, np.abs(gw).max()), (k, xk, gg, gw)


def _compat(g, **kw):
    from generalized_icp_b200 import compat
    return compat.gicp_extended(g["src"], g["tgt"], max_iterations=int(g["max_iterations"]),
                                tolerance=float(g["tolerance"]), max_distance_correspondence=float(g["d_max"]),
                                max_distance_nearest_neighbors=float(g["r_knn"]), **kw)


# Where the engine's end-to-end result is known to leave the reference's: a converged inner solve (the engine, and
# the oracle with inner="newton") against fmin_cg stopping unconverged (warnflag 1/2 on about half of config 1's
# outer iterations).  Established on the CPU with the oracle (tests/golden/make_match_rate.py prints the
# same for seeds 0-39) and asserted here BOTH ways, so neither a regression nor a silent improvement goes unnoticed.
# value: (iteration count of the converged solve, "count" | "T" = what differs from the reference)
KNOWN_DEVIATIONS = {
    "config1_seed0": (16, "count"),          # reference 14; ends 0.155 px / 6e-4 rad away
    "config1_seed2": (23, "count"),          # reference 17
    "config1_seed5": (18, "count"),          # reference 17
    "config2_rays360_pair0": (4, "T"),       # same count; the reference's CG stops 1.4e-3 rad / 0.42 px short
    "config2_rays360_pair4": (5, "count"),   # reference 4
}


@pytest.mark.parametrize("name", SMALL)
def test_end_to_end_vs_reference(name, golden, torch_cuda):
    """Final transform and iteration count against the UNMODIFIED reference's own run (gicp.py:78-174), every
    fixture: identical iteration count and T within 1e-4 rad / 1e-2 px (fmin_cg's own gtol = 1e-5 exit leaves it
    that far from the minimiser), except the fixtures of KNOWN_DEVIATIONS, which must deviate exactly as recorded."""
    g = golden(name)
    r = _compat(g)
    n_ref = len(g["it_fopt"])
    d_ang = abs(np.arctan2(r["T"][1, 0], r["T"][0, 0]) - np.arctan2(g["T"][1, 0], g["T"][0, 0]))
    d_t = np.abs(r["T"][:2, 2] - g["T"][:2, 2]).max()
    if name in KNOWN_DEVIATIONS:
        n_conv, what = KNOWN_DEVIATIONS[name]
        assert r["n_outer"] == n_conv
        if what == "count":
            assert n_conv != n_ref
        else:
            assert n_conv == n_ref and (d_ang >= 1e-4 or d_t >= 1e-2)
    else:
        assert r["n_outer"] == n_ref
        assert len(r["all_T"]) == len(g["all_T"])
        assert len(r["all_src_cov"]) == len(g["all_src_cov"])
        assert len(r["hw_src"]) == int(g["n_hw"])
        assert d_ang < 1e-4 and d_t < 1e-2


def test_match_rate_vs_reference(torch_cuda):
    """The match rate over the BASELINE configs 1-2 against the unmodified reference (outcomes recorded by
    tests/golden/make_match_rate.py): config 1 seeds 0-39, config 2 drives of 30 scans at 90 and 360 rays.
    Reported to gpurun_out/parity_vs_reference.json (committed as profiles/parity_vs_reference_r02.md) and
    asserted not to regress.  The engine must also agree with the converged-inner-solve oracle on every case."""
    import json
    import os
    import demo_inputs
    from conftest import GOLDEN, ROOT
    from generalized_icp_b200 import compat
    m = dict(np.load(os.path.join(GOLDEN, "match_rate_reference.npz")))
    drives = {rays: demo_inputs.lidar_sequence(seed=1, num_rays=rays, n_scans=30)[0] for rays in (90, 360)}
    rows = {}
    for i in range(len(m["kind"])):
        kind, seed, rays, pair = (int(m[k][i]) for k in ("kind", "seed", "rays", "pair"))
        if kind == 1:
            s, t = demo_inputs.config1_pair(seed)
            kw, key = {}, "config1"
        else:
            s, t = np.asarray(drives[rays][pair]), np.asarray(drives[rays][pair + 1])
            kw, key = dict(max_distance_nearest_neighbors=200, tolerance=1), f"config2_rays{rays}"
        r = compat.gicp_extended(s, t, full_history=False, **kw)

        def close(T, tol_ang, tol_t):
            d_ang = abs(np.arctan2(r["T"][1, 0], r["T"][0, 0]) - np.arctan2(T[1, 0], T[0, 0]))
            return bool(d_ang < tol_ang and np.abs(r["T"][:2, 2] - T[:2, 2]).max() < tol_t)

        row = rows.setdefault(key, dict(cases=0, same_iterations_as_reference=0, T_close_to_reference=0,  # 1e-5 rad, 1e-2 px
                                        reference_cg_always_converged=0, same_iterations_as_converged_oracle=0,
                                        T_close_to_converged_oracle=0))   # max |dT| < 1e-6
        row["cases"] += 1
        row["same_iterations_as_reference"] += int(r["n_outer"] == int(m["n_ref"][i]))
        row["T_close_to_reference"] += int(close(m["T_ref"][i], 1e-5, 1e-2))
        row["reference_cg_always_converged"] += int(bool(m["clean"][i]))
        row["same_iterations_as_converged_oracle"] += int(r["n_outer"] == int(m["n_newton"][i]))
        row["T_close_to_converged_oracle"] += int(np.abs(r["T"] - m["T_newton"][i]).max() < 1e-6)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_vs_reference.json"), "w") as f:
        json.dump(dict(rows=rows, versions=str(m["versions"])), f, indent=1)
    # floors = what the converged-inner-solve oracle achieves on the CPU (22/40 and 34/40; 29/29 and 29/29; 28/29 and 26/29)
    c1, c90, c360 = rows["config1"], rows["config2_rays90"], rows["config2_rays360"]
    assert c1["same_iterations_as_reference"] >= 21 and c1["T_close_to_reference"] >= 33
    assert c90["same_iterations_as_reference"] >= 29 and c90["T_close_to_reference"] >= 29
    assert c360["same_iterations_as_reference"] >= 28 and c360["T_close_to_reference"] >= 26
    # against the well-defined target every case must agree; config 1's chaotic trajectories (60 degrees off at the
    # start) are allowed two count mismatches out of 40
    assert c1["same_iterations_as_converged_oracle"] >= 38
    assert c90["same_iterations_as_converged_oracle"] == 29 and c360["same_iterations_as_converged_oracle"] == 29
    assert c90["T_close_to_converged_oracle"] == 29 and c360["T_close_to_converged_oracle"] >= 28


@pytest.mark.parametrize("name", SMALL)
def test_end_to_end_vs_oracle_converged_inner(name, golden, torch_cuda):
    """Against the oracle with a converged inner solve (the well-defined target): identical
    iteration count, T history to 1e-5, final T to 1e-6 (north_star asks 1e-5 rad, 1e-5 * extent)."""
    from oracle import gicp_oracle as O
    g = golden(name)
    r = _compat(g)
    ref = O.gicp_oracle(g["src"], g["tgt"], max_iterations=int(g["max_iterations"]), tolerance=float(g["tolerance"]),
                        max_distance_correspondence=float(g["d_max"]),
                        max_distance_nearest_neighbors=float(g["r_knn"]), inner="newton", recompute_src_cov=True)
    assert r["n_outer"] == ref["n_outer"]
    assert r["converged_at"] == (ref["converged_at"] if ref["converged_at"] is not None else -1)
    # history: far from the solution (config 1 starts 60 deg / 160 px away) the inner problem is
    # ill-conditioned, the two converged solvers agree to ~1e-10 relative (|t| ~ 500 px)
    assert np.abs(np.stack(r["all_T"]) - np.stack(ref["all_T"])).max() < 1e-5
    assert np.abs(r["T"] - ref["T"]).max() < 1e-6
    assert np.abs(np.stack(r["all_src_cov"]) - np.stack(ref["all_src_cov"])).max() < 1e-5
    assert len(r["hw_src"]) == len(ref["hw_src"])
    for a, b in zip(r["hw_src"], ref["hw_src"]):
        assert a.shape == b.shape


def test_compat_tuple_and_print(golden, torch_cuda, capsys):
    import gicp as shim
    g = golden("config2_rays90_pair0")
    src = [tuple(p) for p in g["src"]]      # lists of tuples, robot-visualization.py:155-156
    tgt = [tuple(p) for p in g["tgt"]]
    out = shim.gicp(src, tgt, max_distance_nearest_neighbors=200, tolerance=1)
    assert len(out) == 7
    T, all_T, c0, ct, hs, ht, allc = out
    assert T.shape == (3, 3) and T.dtype == np.float64 and T is not None
    assert np.array_equal(all_T[0], np.eye(3)) and np.array_equal(T, all_T[-1])
    assert c0.shape == (len(src), 2, 2) and ct.shape == (len(tgt), 2, 2)
    assert np.array_equal(allc[0], c0)
    # gicp.py:170-172: the five highest-weight correspondences of the first iteration (T = I: independent of the
    # inner solver), recorded from the reference; every iteration's set comes out of ONE batched call
    assert len(hs) == len(ht) == len(all_T) - 1
    assert np.abs(hs[0] - g["hw_src_0"]).max() < 1e-9 and np.abs(ht[0] - g["hw_tgt_0"]).max() < 1e-9
    assert "Converged at iteration" in capsys.readouterr().out
    import pickle
    pickle.loads(pickle.dumps(out))


# ------------------------------------------------------------------------------------------------
# 3-D (no reference implementation exists: parity is against the oracle's 3-D restatement)
# ------------------------------------------------------------------------------------------------
P3 = dict(k=20, max_distance_nearest_neighbors=4.0, max_distance_correspondence=2.0)


def _pair3(seed, n=3000):
    from generalized_icp_b200 import synthetic
    return synthetic.patches3d_pair(n=n, n_patches=5, cube=30.0, patch=20.0, seed=seed)


@pytest.mark.parametrize("storage", ["f32", "f64"])
def test_3d_stages_vs_oracle(storage, torch_cuda):
    torch = torch_cuda
    from generalized_icp_b200.engine import GicpEngine
    from oracle import gicp_oracle as O
    src, tgt, Tgt = _pair3(0)
    eng = GicpEngine(3, storage)
    eng.set_params(**P3, tolerance=1e-6)
    dt = torch.float32 if storage == "f32" else torch.float64
    eng.set_target(torch.as_tensor(tgt, dtype=dt, device="cuda"))
    eng.set_source(torch.as_tensor(src, dtype=dt, device="cuda"))
    s64, t64 = src.astype(np.float64), tgt.astype(np.float64)
    for which, cloud in ((0, s64), (1, t64)):
        want, _ = O.knn_bruteforce(cloud, 20, 4.0)
        want = np.where(want == len(cloud), -1, want)
        idx, _ = eng.knn(which)
        assert np.array_equal(idx.cpu().numpy(), want)          # bit-exact, also with fp32 storage
        cov_want = O.covariances_from_neighbors(cloud, np.where(want < 0, len(cloud), want))
        tol = 2e-4 if storage == "f32" else 1e-7                 # fp32 storage rounds C (entries <= 100) to 6e-6
        assert np.abs(eng.covariances(which).cpu().numpy() - cov_want).max() < tol
    # correspondences at identity and at the ground truth
    for T in (np.eye(4), Tgt):
        moved = O.apply_transformation(s64, T)
        want_idx, want_d = O.correspond(moved, t64, 2.0, method="brute")
        idx, dist, W = eng.correspond(T)
        assert np.array_equal(idx.cpu().numpy(), want_idx)
        m = want_idx >= 0
        assert np.abs(dist.cpu().numpy()[m] - want_d[m]).max() < 1e-9
        cs = eng.covariances(0).cpu().numpy()
        ct = eng.covariances(1).cpu().numpy()
        Wwant = O.weights(T[:3, :3] @ cs @ T[:3, :3].T, ct, want_idx)
        assert np.abs(W.cpu().numpy() - Wwant).max() < 1e-9
        # reduced form reproduces the loss at a nearby transform
        from generalized_icp_b200.engine import reduced_form_loss
        red = eng.normal_equations(T)[0]
        x = np.concatenate([T[:3, 3] + [0.01, -0.02, 0.005], O.rotvec_from_matrix(T[:3, :3]) + [1e-3, -2e-3, 5e-4]])
        Te = O.offset_to_matrix(x, 3)
        want_loss = O.loss(x, s64, O.corresponding_points(t64, want_idx), Wwant)
        got_loss = reduced_form_loss(red, 3, T, Te)
        rel = 2e-5 if storage == "f32" else 1e-9
        assert abs(got_loss - want_loss) <= rel * abs(want_loss), (got_loss, want_loss)


@pytest.mark.parametrize("storage,seed", [("f64", 0), ("f64", 1), ("f32", 0), ("f32", 1), ("f32", 2)])
def test_3d_end_to_end_vs_oracle(storage, seed, torch_cuda):
    """north_star tolerance: identical iteration count, rotation angle <= 1e-5 rad,
    translation <= 1e-5 * extent (extent 30 m here)."""
    from generalized_icp_b200 import compat
    from oracle import gicp_oracle as O
    src, tgt, _ = _pair3(seed)
    r = compat.gicp_extended(src, tgt, storage=storage, full_history=False, **P3)
    ref = O.gicp_oracle(src, tgt, inner="newton", recompute_src_cov=True, record=False, **P3)
    assert r["n_outer"] == ref["n_outer"]
    dR = r["T"][:3, :3].T @ ref["T"][:3, :3]
    ang = np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1))
    assert ang <= 1e-5
    assert np.linalg.norm(r["T"][:3, 3] - ref["T"][:3, 3]) <= 1e-5 * 30.0


def test_batch_equals_single(torch_cuda):
    """A batch of pairs gives bit-identical results to running each pair alone (independent pairs,
    no cross-talk, deterministic reductions)."""
    torch = torch_cuda
    from generalized_icp_b200.engine import GicpEngine
    pairs = [_pair3(s, n=1500 + 137 * s) for s in range(4)]
    eng = GicpEngine(3, "f32")
    eng.set_params(**P3)
    singles = []
    for s, t, _ in pairs:
        eng.set_target(torch.as_tensor(t, device="cuda"))
        eng.set_source(torch.as_tensor(s, device="cuda"))
        r = eng.register()
        singles.append((r.T[0].cpu().numpy(), int(r.n_outer[0])))
    S = torch.as_tensor(np.concatenate([p[0] for p in pairs]), device="cuda")
    T = torch.as_tensor(np.concatenate([p[1] for p in pairs]), device="cuda")
    off = np.concatenate([[0], np.cumsum([len(p[0]) for p in pairs])])
    eng.set_target(T, off)
    eng.set_source(S, off)
    r = eng.register()
    for i, (Ti, ni) in enumerate(singles):
        assert int(r.n_outer[i]) == ni
        assert np.array_equal(r.T[i].cpu().numpy(), Ti)
    r2 = eng.register()                                           # run-to-run bitwise reproducibility
    assert torch.equal(r.T, r2.T) and torch.equal(r.loss_hist.nan_to_num(), r2.loss_hist.nan_to_num())


def test_edge_cases(torch_cuda):
    torch = torch_cuda
    from generalized_icp_b200 import compat
    # tiny clouds, isolated points (identity covariance, gicp.py:33-34), everything gated out
    src = np.array([[0.0, 0.0], [1.0, 0.5], [2.0, 0.1], [500.0, 500.0]])
    tgt = src + np.array([0.3, -0.2])
    r = compat.gicp_extended(src, tgt, max_distance_nearest_neighbors=5.0, max_distance_correspondence=10.0)
    assert np.array_equal(r["src_cov0"][3], np.eye(2))
    assert np.isfinite(r["T"]).all()
    far = compat.gicp_extended(src, tgt + 1e4, max_distance_nearest_neighbors=5.0, max_distance_correspondence=10.0)
    assert np.array_equal(far["T"], np.eye(3)) and far["converged_at"] == 1   # loss 0 twice -> stop at it 1
    one = compat.gicp_extended(src[:1], tgt[:1])
    assert np.array_equal(one["src_cov0"][0], np.eye(2))
    with pytest.raises(Exception):
        compat.gicp_extended(np.zeros((3, 4)), np.zeros((3, 4)))


# ------------------------------------------------------------------------------------------------
# streaming scan-sequence odometry (grid + covariance reuse)
# ------------------------------------------------------------------------------------------------
def test_scan_sequence_reuse_is_bit_identical(torch_cuda):
    """Promoting the target to the next pair's source gives bit-identical transforms to registering
    every pair from scratch through the drop-in gicp(), and the integrated pose follows the
    simulated robot (robot-visualization.py:258-265)."""
    import demo_inputs
    from generalized_icp_b200 import compat
    from generalized_icp_b200.odometry import ScanOdometry, integrate_pose, trajectory_errors
    scans, truth = demo_inputs.lidar_sequence(seed=2, num_rays=360, n_scans=14)
    odo = ScanOdometry(start_pose=(50.0, 400.0, 0.0))
    for s in scans:
        odo.push(s)
    assert len(odo.transforms) == len(scans) - 1
    pose = (50.0, 400.0, 0.0)
    for i in range(len(scans) - 1):
        r = compat.gicp_extended(np.asarray(scans[i]), np.asarray(scans[i + 1]), max_distance_nearest_neighbors=200,
                                 tolerance=1, full_history=False)
        assert np.array_equal(r["T"], odo.transforms[i])
        assert r["n_outer"] == odo.iterations[i]
        pose = integrate_pose(pose, r["T"])
    assert np.allclose(pose, odo.pose)
    # the first scan is taken after 5 ticks of driving: compare displacements from the first pose
    est = np.asarray(odo.poses)
    tru = np.asarray(truth)
    est_rel = est - est[0]
    tru_rel = tru - tru[0]
    err = trajectory_errors(est_rel + [0, 0, 0], np.column_stack([tru_rel[:, 0], tru_rel[:, 1], tru_rel[:, 2]]))
    assert err["position_max"] < 25.0          # px, after 13 pairs of ~10 px steps with +-2 px range noise
    assert err["orientation_max"] < 0.2


# ------------------------------------------------------------------------------------------------
# k-NN corner cases: every selection path (fast list path, overflow hand-over, general register path)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,storage,k,radius,n", [
    (3, "f32", 1, 2.0, 2000),      # k = 1: only the point itself
    (3, "f32", 3, 1.0, 2000),      # small radius: many points with fewer than k neighbours
    (3, "f32", 24, 4.0, 3000),     # largest k of the fast path
    (3, "f32", 32, 6.0, 3000),     # general path (k > 24)
    (3, "f64", 20, 4.0, 2000),     # float64 storage with k = 20 -> general path
    (2, "f32", 6, 60.0, 400),      # 2-D, float32 storage
    (3, "f32", 6, 400.0, 3000),    # radius >> spacing: all k in histogram bin 0 -> list overflow -> hand-over
    (2, "f64", 6, 1e6, 300),       # same in 2-D float64
])
def test_knn_paths_bit_exact(dim, storage, k, radius, n, torch_cuda):
    torch = torch_cuda
    from generalized_icp_b200.engine import GicpEngine
    from oracle import gicp_oracle as O
    rng = np.random.default_rng(k * 1000 + n + dim)
    if dim == 3:
        from generalized_icp_b200 import synthetic
        pts, _, _ = synthetic.patches3d_pair(n=n, n_patches=4, cube=30.0, patch=20.0, seed=k)
    else:
        pts = rng.uniform(0, 800, size=(n, 2))
    dt = torch.float32 if storage == "f32" else torch.float64
    dev_pts = torch.as_tensor(pts, device="cuda").to(dt)
    cloud = dev_pts.double().cpu().numpy()            # what the engine sees (fp32-rounded for f32 storage)
    eng = GicpEngine(dim, storage)
    eng.set_params(k=k, max_distance_nearest_neighbors=radius, max_distance_correspondence=radius)
    eng.set_target(dev_pts)
    idx, dist = eng.knn(1)
    want, wantd = O.knn_bruteforce(cloud, k, radius)
    want = np.where(want == len(cloud), -1, want)
    assert np.array_equal(idx.cpu().numpy(), want)
    d = dist.cpu().numpy()
    m = want >= 0
    assert np.abs(d[m] - wantd[m]).max() < 1e-9 * max(1.0, radius)
    cov = eng.covariances(1).cpu().numpy()
    cov_want = O.covariances_from_neighbors(cloud, np.where(want < 0, len(cloud), want))
    # near-isotropic neighbourhoods have an ill-defined normal: compare where the oracle's own
    # eigen-gap is healthy, and always require a valid covariance (symmetric, eigenvalues in {1, 10, 100})
    ev = np.linalg.eigvalsh(cov)
    assert np.isfinite(cov).all() and ev.min() > 0.99 and ev.max() < 100.01
    close = np.abs(cov - cov_want).reshape(len(cloud), -1).max(1) < (2e-3 if storage == "f32" else 1e-6)
    # eigen-gap of the neighbourhood scatter (2 points, collinear points: the normal is not unique)
    nb = np.where(want < 0, 0, want)
    w = (want >= 0)[..., None].astype(float)
    cnt = (want >= 0).sum(1)
    mean = (cloud[nb] * w).sum(1) / np.maximum(cnt, 1)[:, None]
    dev = (cloud[nb] - mean[:, None, :]) * w
    lam = np.linalg.eigvalsh(np.einsum("nki,nkj->nij", dev, dev))
    healthy = (cnt >= dim) & ((lam[:, 1] - lam[:, 0]) > 1e-2 * lam[:, -1])
    if healthy.any():
        assert close[healthy].mean() > 0.995
    assert np.array_equal(cov[cnt <= 1], np.broadcast_to(np.eye(dim), (int((cnt <= 1).sum()), dim, dim)))


def test_knn_overflow_many_chunks_per_warp(torch_cuda):
    """More overflow chunks than the hand-over kernel has warps (296 x 4): every warp of the general
    path then serves several chunks in a row on the same TMA stage and mbarrier."""
    torch = torch_cuda
    from generalized_icp_b200 import synthetic
    from generalized_icp_b200.engine import GicpEngine
    from oracle import gicp_oracle as O
    n, k, radius = 48_000, 6, 400.0     # 1500 chunks, all of them overflow (radius >> spacing)
    pts, _, _ = synthetic.patches3d_pair(n=n, n_patches=8, cube=60.0, patch=30.0, seed=11)
    dev_pts = torch.as_tensor(pts, device="cuda").float()
    cloud = dev_pts.double().cpu().numpy()
    eng = GicpEngine(3, "f32")
    eng.set_params(k=k, max_distance_nearest_neighbors=radius, max_distance_correspondence=radius)
    eng.set_target(dev_pts)
    idx, dist = eng.knn(1)
    want, wantd = O.knn_kdtree(cloud, k, radius)
    got = idx.cpu().numpy()
    # cKDTree breaks exact distance ties arbitrarily: compare the distances everywhere, the indices
    # where the neighbour distances are distinct
    d = dist.cpu().numpy()
    assert np.abs(d - wantd).max() < 1e-9 * radius
    distinct = (np.diff(wantd, axis=1) > 0).all(1)
    assert distinct.mean() > 0.99
    assert np.array_equal(got[distinct], want[distinct])
    cov = eng.covariances(1).cpu().numpy()
    ev = np.linalg.eigvalsh(cov)
    assert np.isfinite(cov).all() and ev.min() > 0.99 and ev.max() < 100.01


def test_duplicate_points_tie_break(torch_cuda):
    """Exact ties (duplicated points) go to the lower index, like the canonical oracle rule."""
    torch = torch_cuda
    from generalized_icp_b200.engine import GicpEngine
    from oracle import gicp_oracle as O
    rng = np.random.default_rng(3)
    base = rng.uniform(0, 100, size=(60, 2))
    pts = np.concatenate([base, base[:30], base[:10]])      # exact duplicates
    rng.shuffle(pts)
    eng = GicpEngine(2, "f64")
    eng.set_params(k=6, max_distance_nearest_neighbors=40.0)
    eng.set_target(torch.as_tensor(pts, device="cuda"))
    idx, _ = eng.knn(1)
    want, _ = O.knn_bruteforce(pts, 6, 40.0)
    assert np.array_equal(idx.cpu().numpy(), np.where(want == len(pts), -1, want))
    # 1-NN of the cloud against itself under the identity: every point matches the LOWEST index among its copies
    eng.set_params(k=6, max_distance_nearest_neighbors=40.0, max_distance_correspondence=5.0)
    eng.set_target(torch.as_tensor(pts, device="cuda"))
    eng.set_source(torch.as_tensor(pts, device="cuda"))
    nn, _, _ = eng.correspond(np.eye(3), with_W=False)
    want_nn, _ = O.correspond(pts, pts, 5.0, method="brute")
    assert np.array_equal(nn.cpu().numpy(), want_nn)


# ------------------------------------------------------------------------------------------------
# ICP-family variants through the same kernels (only the covariance model changes)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model", [1, 2])
def test_icp_variants_vs_oracle(model, torch_cuda):
    from generalized_icp_b200 import compat
    from oracle import gicp_oracle as O
    src, tgt, _ = _pair3(4)
    r = compat.gicp_extended(src, tgt, storage="f64", full_history=False, covariance_model=model, **P3)
    ref = O.gicp_oracle(src, tgt, inner="newton", record=False, covariance_model=model, **P3)
    assert r["n_outer"] == ref["n_outer"]
    dR = r["T"][:3, :3].T @ ref["T"][:3, :3]
    assert np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1)) <= 1e-5
    assert np.linalg.norm(r["T"][:3, 3] - ref["T"][:3, 3]) <= 1e-5 * 30.0
    if model == 1:
        assert np.array_equal(r["tgt_cov"][0], np.eye(3)) and not r["src_cov0"].any()


# ------------------------------------------------------------------------------------------------
# GPU LiDAR ray caster vs the headless restatement of robot-visualization.py:42-120
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("num_rays", [90, 360])
def test_ray_caster_matches_demo(num_rays, torch_cuda):
    import math
    import random
    import demo_inputs
    from generalized_icp_b200.engine import ray_cast
    sim = demo_inputs.LidarSim(seed=0, num_rays=num_rays)
    poses, want, noises = [], [], []
    for step in range(25):
        sim.step(up=True, right=(step % 7 == 3), left=(step % 11 == 5))
        poses.append((sim.x, sim.y, sim.yaw))
        # replay cast_ray with a recorded noise stream so the device gets the same numbers
        rec = []
        class Rec(random.Random):
            def uniform(self, a, b):
                v = super().uniform(a, b)
                rec.append(v)
                return v
        sim.rnd = Rec(1000 + step)
        rows, nz = [], []
        for angle in range(sim.yaw, sim.yaw + 360, 360 // num_rays):
            before = len(rec)
            d = sim.cast_ray(angle)
            nz.append(rec[-1] if len(rec) > before else 0.0)
            if d:
                rows.append((d * math.cos(math.radians(angle - sim.yaw)), d * math.sin(math.radians(angle - sim.yaw))))
            else:
                rows.append(None)
        want.append(rows)
        noises.append(nz)
    rel, hit = ray_cast(np.asarray(poses, dtype=np.float64), num_rays=num_rays, noise=np.asarray(noises))
    rel, hit = rel.cpu().numpy(), hit.cpu().numpy()
    for i, rows in enumerate(want):
        for j, r in enumerate(rows):
            assert bool(hit[i, j]) == (r is not None), (i, j)
            if r is not None:
                assert abs(rel[i, j, 0] - r[0]) < 1e-9 and abs(rel[i, j, 1] - r[1]) < 1e-9


def test_register_host_batch_matches_device_batch(torch_cuda):
    """The chunked, copy-overlapped host-batch path returns exactly what one device-resident batch returns."""
    torch = torch_cuda
    from generalized_icp_b200.engine import GicpEngine
    pairs = [_pair3(s, n=1200 + 100 * s) for s in range(5)]
    S = np.concatenate([p[0] for p in pairs])
    T = np.concatenate([p[1] for p in pairs])
    off = np.concatenate([[0], np.cumsum([len(p[0]) for p in pairs])])
    eng = GicpEngine(3, "f32")
    eng.set_params(**P3)
    eng.set_target(torch.as_tensor(T, device="cuda"), off)
    eng.set_source(torch.as_tensor(S, device="cuda"), off)
    r = eng.register()
    hS, hT = torch.as_tensor(S).pin_memory(), torch.as_tensor(T).pin_memory()
    T_h, n_h, c_h = eng.register_host_batch(hS, hT, off, chunk_pairs=2)
    assert np.array_equal(T_h.numpy(), r.T.cpu().numpy())
    assert np.array_equal(n_h.numpy(), r.n_outer.cpu().numpy())
    assert np.array_equal(c_h.numpy(), r.converged_at.cpu().numpy())


# ------------------------------------------------------------------------------------------------
# round-2 review items
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model", [1, 2])
def test_scan_sequence_reuse_with_icp_variants(model, torch_cuda):
    """gicpPromoteTargetToSource with covariance_model 1 / 2: the promoted side held TARGET covariances
    (I, or the estimated ones); as a source it must have C = 0 like a freshly set source."""
    import demo_inputs
    from generalized_icp_b200 import compat
    from generalized_icp_b200.odometry import ScanOdometry
    scans, _ = demo_inputs.lidar_sequence(seed=4, num_rays=360, n_scans=6)
    odo = ScanOdometry(start_pose=(50.0, 400.0, 0.0), covariance_model=model)
    for s in scans:
        odo.push(s)
    for i in range(len(scans) - 1):
        r = compat.gicp_extended(np.asarray(scans[i]), np.asarray(scans[i + 1]), max_distance_nearest_neighbors=200,
                                 tolerance=1, full_history=False, covariance_model=model)
        assert np.array_equal(r["T"], odo.transforms[i]), i
        assert r["n_outer"] == odo.iterations[i]


def test_empty_clouds_inside_a_batch(torch_cuda):
    """An empty source cloud FIRST in the batch (a LiDAR scan with zero hits), an empty target cloud, and a
    regular pair after them: no out-of-bounds read, identity for the empty pairs, the regular pair unchanged."""
    torch = torch_cuda
    from generalized_icp_b200.engine import GicpEngine
    sA, tA, _ = _pair3(0, n=1400)
    sB, tB, _ = _pair3(1, n=1500)
    eng = GicpEngine(3, "f32")
    eng.set_params(**P3)
    eng.set_target(torch.as_tensor(tB, device="cuda"))
    eng.set_source(torch.as_tensor(sB, device="cuda"))
    alone = eng.register()
    S = torch.as_tensor(np.concatenate([sA, sB]), device="cuda")
    T = torch.as_tensor(np.concatenate([tA, tB]), device="cuda")
    eng.set_target(T, [0, len(tA), len(tA), len(tA) + len(tB)])          # pair 1: empty target
    eng.set_source(S, [0, 0, len(sA), len(sA) + len(sB)])                # pair 0: empty source
    r = eng.register()
    torch.cuda.synchronize()
    eye = np.eye(4)
    assert np.array_equal(r.T[0].cpu().numpy(), eye) and np.array_equal(r.T[1].cpu().numpy(), eye)
    assert int(r.converged_at[0]) == 1 and int(r.converged_at[1]) == 1   # loss 0 twice -> stop at iteration 1
    assert np.array_equal(r.T[2].cpu().numpy(), alone.T[0].cpu().numpy())
    assert int(r.n_outer[2]) == int(alone.n_outer[0])
    idx, _, _ = eng.correspond(np.stack([eye, eye, eye]), with_W=False)
    assert (idx[:len(sA)].cpu().numpy() == -1).all()                      # nothing to match in an empty target
