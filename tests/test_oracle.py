"""The CPU oracle against the fixtures recorded from the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import golden_names
from oracle import gicp_oracle as O

ALL = golden_names()
SMALL = [n for n in ALL if n != "config1_seed4"]


@pytest.mark.parametrize("name", ALL)
def test_knn_sets_bit_exact(name, golden):
    g = golden(name)
    for cloud, key in ((g["src"], "src_knn"), (g["tgt"], "tgt_knn")):
        bf, _ = O.knn_bruteforce(cloud, 6, float(g["r_knn"]))
        kd, _ = O.knn_kdtree(cloud, 6, float(g["r_knn"]))
        assert np.array_equal(bf, g[key])          # canonical rule == scipy per-point query of the reference
        assert np.array_equal(kd, g[key])


@pytest.mark.parametrize("name", ALL)
def test_covariances(name, golden):
    g = golden(name)
    cs, _ = O.compute_covariances(g["src"], float(g["r_knn"]))
    ct, _ = O.compute_covariances(g["tgt"], float(g["r_knn"]))
    assert np.abs(cs - g["src_cov0"]).max() < 1e-9
    assert np.abs(ct - g["tgt_cov"]).max() < 1e-9
    # gicp.py:120 recomputation == R C0 R^T (SURVEY appendix A rule 5)
    for k, T in enumerate(g["all_T"][: len(g["all_src_cov"])]):
        R = T[:2, :2]
        assert np.abs(R @ cs @ R.T - g["all_src_cov"][k]).max() < 1e-9


@pytest.mark.parametrize("name", SMALL)
def test_correspondences_weights_loss_grad(name, golden):
    g = golden(name)
    src, tgt = g["src"], g["tgt"]
    for k in range(len(g["it_fopt"])):
        T = g["all_T"][k]
        moved = O.apply_transformation(src, T)
        idx, dist = O.correspond(moved, tgt, float(g["d_max"]))
        assert np.array_equal(idx, g["nn_idx"][k])
        assert np.abs(dist - g["nn_dist"][k]).max() < 1e-12
        idx_bf, _ = O.correspond(moved, tgt, float(g["d_max"]), method="brute")
        assert np.array_equal(idx_bf, idx)
        q = O.corresponding_points(tgt, idx)
        assert np.array_equal(q, g["it_q"][k])
        W = O.weights(g["all_src_cov"][k], g["tgt_cov"], idx)
        assert np.abs(W - g["it_W"][k]).max() < 1e-12
        for xk, lk, gk in (("it_xopt", "it_loss_at_xopt", "it_grad_at_xopt"), ("it_x0", "it_loss_at_x0", "it_grad_at_x0")):
            x = g[xk][k]
            l_ref, g_ref = float(g[lk][k]), g[gk][k]
            assert abs(O.loss(x, src, q, W) - l_ref) <= 1e-12 * max(1.0, abs(l_ref))
            assert np.abs(O.grad_loss(x, src, q, W) - g_ref).max() <= 1e-9 * max(1.0, np.abs(g_ref).max())


@pytest.mark.parametrize("name", [n for n in SMALL if n.startswith("config2")] + ["config1_seed0"])
def test_end_to_end_fidelity(name, golden):
    """inner='cg' follows the reference call for call: same iteration count,
    same history (the vectorised arithmetic differs from the reference's
    per-point loops only in rounding)."""
    g = golden(name)
    out = O.gicp_oracle(g["src"], g["tgt"], max_iterations=int(g["max_iterations"]), tolerance=float(g["tolerance"]),
                        max_distance_correspondence=float(g["d_max"]),
                        max_distance_nearest_neighbors=float(g["r_knn"]), inner="cg")
    if (g["it_warnflag"] == 0).all():
        assert len(out["all_T"]) == len(g["all_T"])
        assert np.abs(np.stack(out["all_T"]) - g["all_T"]).max() < 1e-4  # CG stops at |g|inf<=1e-5: rounding moves its exit point
        assert len(out["all_src_cov"]) == len(g["all_src_cov"])
        assert len(out["hw_src"]) == int(g["n_hw"])
    else:
        # fmin_cg did not converge somewhere: the reference trajectory is not a well-defined target
        assert out["n_outer"] >= 1


@pytest.mark.parametrize("name", [n for n in SMALL if n.startswith("config2")])
def test_newton_inner_matches_converged_cg(name, golden):
    g = golden(name)
    src = g["src"]
    for k in range(len(g["it_fopt"])):
        if int(g["it_warnflag"][k]) != 0:
            continue
        q, W = g["it_q"][k], g["it_W"][k]
        x, f, _, _ = O.inner_newton(g["it_x0"][k], src, q, W)
        assert f <= float(g["it_fopt"][k]) + 1e-9
        assert np.abs(O.grad_loss(x, src, q, W)).max() < 1e-6
        # fmin_cg stops at |g|inf <= 1e-5, so it sits within ~1e-5/curvature of the minimiser
        assert abs(x[2] - g["it_xopt"][k][2]) < 1e-5
        assert np.abs(x[:2] - g["it_xopt"][k][:2]).max() < 1e-3


def test_3d_gradient_is_consistent():
    rng = np.random.default_rng(0)
    src = rng.normal(size=(50, 3)) * 3
    q = src + rng.normal(size=(50, 3)) * 0.1
    A = rng.normal(size=(50, 3, 3))
    W = A @ A.transpose(0, 2, 1) + np.eye(3)
    x = np.array([0.1, -0.2, 0.3, 0.2, -0.1, 0.4])
    g = O.grad_loss(x, src, q, W)
    num = np.zeros(6)
    for a in range(6):
        e = np.zeros(6)
        e[a] = 1e-6
        num[a] = (O.loss(x + e, src, q, W) - O.loss(x - e, src, q, W)) / 2e-6
    assert np.abs(g - num).max() < 1e-5 * max(1, np.abs(g).max())
    xs, f, _, _ = O.inner_newton(x, src, q, W)
    assert np.abs(O.grad_loss(xs, src, q, W)).max() < 1e-7
    xc, fc, warn, _ = O.inner_cg(x, src, q, W)
    assert f <= fc + 1e-9


def test_3d_registration_recovers_motion():
    import importlib.util, os, sys
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("synthetic", os.path.join(ROOT, "generalized-icp_b200", "synthetic.py"))
    syn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(syn)
    src, tgt, T = syn.patches3d_pair(n=4000, n_patches=6, cube=30.0, patch=20.0, seed=3)
    out = O.gicp_oracle(src, tgt, k=20, max_distance_nearest_neighbors=4.0, max_distance_correspondence=2.0,
                        inner="newton", recompute_src_cov=False, record=False)
    Te = out["T"]
    ang = np.arccos(np.clip((np.trace(Te[:3, :3].T @ T[:3, :3]) - 1) / 2, -1, 1))
    assert ang < 2e-3
    assert np.linalg.norm(Te[:3, 3] - T[:3, 3]) < 0.05
    assert out["converged_at"] is not None
