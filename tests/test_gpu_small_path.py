"""The latency path for small clouds (single-block grid build `small_grid_kernel`, one-launch k-NN, fused
registration loop `register_loop_kernel`) against the multi-launch path on the same inputs: the same grids
(bit-identical neighbour lists), covariances and transforms.  Every existing parity test of the 2-D fixtures runs
through the latency path by default; this file pins the two paths to each other and counts the launches."""
import os

import numpy as np
import pytest

import demo_inputs
from conftest import golden_names

pytestmark = pytest.mark.gpu
SWITCHES = ("GICP_SMALL_GRID", "GICP_FUSED_LOOP", "GICP_KNN_BRUTE")


@pytest.fixture()
def both_paths():
    saved = {k: os.environ.get(k) for k in SWITCHES}

    def run(fn):
        out = []
        for mode in ("1", "0"):
            for k in SWITCHES:
                os.environ[k] = mode          # the library reads the switches at every call
            out.append(fn())
        return out

    yield run
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def _run2d(src, tgt, **kw):
    import torch
    from generalized_icp_b200.engine import GicpEngine
    eng = GicpEngine(2, "f64")
    eng.set_params(k=6, **kw)
    l0 = eng.launch_count
    eng.set_target(torch.as_tensor(tgt, dtype=torch.float64, device="cuda"))
    eng.set_source(torch.as_tensor(src, dtype=torch.float64, device="cuda"))
    r = eng.register(history=True)
    n = int(r.n_outer[0])
    launches = eng.launch_count - l0
    return dict(T=r.T[0].cpu().numpy(), n=n, T_hist=r.T_hist[0, :n].cpu().numpy(), loss=r.loss_hist[0, :n].cpu().numpy(),
                knn_s=eng.knn(0)[0].cpu().numpy(), knn_t=eng.knn(1)[0].cpu().numpy(),
                cov_s=eng.covariances(0).cpu().numpy(), cov_t=eng.covariances(1).cpu().numpy(), launches=launches)


def _same(a, b, tol_T=1e-9, tol_cov=1e-9):
    assert a["n"] == b["n"]
    assert np.array_equal(a["knn_s"], b["knn_s"]) and np.array_equal(a["knn_t"], b["knn_t"])
    assert np.abs(a["cov_s"] - b["cov_s"]).max() < tol_cov and np.abs(a["cov_t"] - b["cov_t"]).max() < tol_cov
    assert np.abs(a["T_hist"] - b["T_hist"]).max() < tol_T
    assert np.abs(a["T"] - b["T"]).max() < tol_T
    assert np.allclose(a["loss"], b["loss"], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("rays", [90, 360])
def test_scan_pairs_both_paths(rays, both_paths):
    scans, _ = demo_inputs.lidar_sequence(seed=2, num_rays=rays, n_scans=4)
    for i in range(3):
        s, t = np.asarray(scans[i], dtype=np.float64), np.asarray(scans[i + 1], dtype=np.float64)
        fused, multi = both_paths(lambda: _run2d(s, t, max_distance_nearest_neighbors=200.0, tolerance=1.0))
        _same(fused, multi)
        # one launch each for the grid build and the k-NN of either cloud, state initialisation + the fused loop
        assert fused["launches"] <= 8, fused["launches"]
        assert multi["launches"] > 3 * fused["launches"]


@pytest.mark.parametrize("name", [n for n in golden_names() if n.startswith("config1")][:3])
def test_config1_both_paths(name, golden, both_paths):
    g = golden(name)
    fused, multi = both_paths(lambda: _run2d(g["src"], g["tgt"], max_iterations=int(g["max_iterations"]),
                                             tolerance=float(g["tolerance"]), max_distance_correspondence=float(g["d_max"]),
                                             max_distance_nearest_neighbors=float(g["r_knn"])))
    _same(fused, multi, tol_T=1e-7)


def test_batch_of_small_pairs_and_empty_cloud(both_paths):
    """29 scan pairs in one batch (one block per pair in every stage) + an empty cloud in the middle."""
    import torch
    from generalized_icp_b200.engine import GicpEngine
    scans, _ = demo_inputs.lidar_sequence(seed=5, num_rays=180, n_scans=12)
    src = [np.asarray(s, dtype=np.float64) for s in scans[:-1]]
    tgt = [np.asarray(s, dtype=np.float64) for s in scans[1:]]
    src[4] = np.zeros((0, 2))

    def run():
        eng = GicpEngine(2, "f64")
        eng.set_params(k=6, max_distance_nearest_neighbors=200.0, tolerance=1.0)
        off_s = np.concatenate([[0], np.cumsum([len(a) for a in src])])
        off_t = np.concatenate([[0], np.cumsum([len(a) for a in tgt])])
        eng.set_target(torch.as_tensor(np.concatenate(tgt), device="cuda"), off_t)
        eng.set_source(torch.as_tensor(np.concatenate(src), device="cuda"), off_s)
        r = eng.register(history=False)
        return r.T.cpu().numpy(), r.n_outer.cpu().numpy(), eng.knn(1)[0].cpu().numpy()

    (Tf, nf, kf), (Tm, nm, km) = both_paths(run)
    assert np.array_equal(nf, nm) and np.array_equal(kf, km)
    assert np.abs(Tf - Tm).max() < 1e-9
    # and each pair of the batch equals the pair registered alone
    one = _run2d(src[7], tgt[7], max_distance_nearest_neighbors=200.0, tolerance=1.0)
    assert one["n"] == nf[7] and np.abs(one["T"] - Tf[7]).max() < 1e-12


def test_small_3d_f32_both_paths(both_paths):
    import torch
    from generalized_icp_b200 import synthetic
    from generalized_icp_b200.engine import GicpEngine
    s3, t3, _ = synthetic.patches3d_pair(n=1800, n_patches=4, cube=20.0, patch=15.0, seed=3)

    def run():
        eng = GicpEngine(3, "f32")
        eng.set_params(k=20, max_distance_nearest_neighbors=4.0, max_distance_correspondence=2.0)
        eng.set_target(torch.as_tensor(t3, device="cuda"))
        eng.set_source(torch.as_tensor(s3, device="cuda"))
        r = eng.register(history=False)
        return r.T[0].cpu().numpy(), int(r.n_outer[0]), eng.knn(0)[0].cpu().numpy(), eng.covariances(1).cpu().numpy()

    (Tf, nf, kf, cf), (Tm, nm, km, cm) = both_paths(run)
    assert nf == nm and np.array_equal(kf, km)
    assert np.abs(cf - cm).max() < 2e-4
    assert np.abs(Tf - Tm).max() < 1e-5


def test_set_pair_equals_separate_calls():
    """gicpSetPair (both sides in one call; the two single-block set-ups run concurrently on the device) against
    gicpSetTarget + gicpSetSource: bit-identical grids, covariances and registration, repeated so that a race between
    the two internal streams would show; larger pairs (multi-kernel set-up) side by side and one after the other."""
    import torch
    from generalized_icp_b200 import synthetic
    from generalized_icp_b200.engine import GicpEngine
    scans, _ = demo_inputs.lidar_sequence(seed=5, num_rays=360, n_scans=5)
    eng_a, eng_b = GicpEngine(2, "f64"), GicpEngine(2, "f64")
    for e in (eng_a, eng_b):
        e.set_params(k=6, max_distance_nearest_neighbors=200.0, tolerance=1.0)
    for rep in range(3):
        for i in range(4):
            s = torch.as_tensor(np.asarray(scans[i]), dtype=torch.float64, device="cuda")
            t = torch.as_tensor(np.asarray(scans[i + 1]), dtype=torch.float64, device="cuda")
            eng_a.set_target(t)
            eng_a.set_source(s)
            ra = eng_a.register(history=True)
            eng_b.set_pair(t, s)
            rb = eng_b.register(history=True)
            assert int(ra.n_outer[0]) == int(rb.n_outer[0])
            assert torch.equal(ra.T, rb.T)
            n = int(ra.n_outer[0])
            assert torch.equal(ra.T_hist[0, :n], rb.T_hist[0, :n])
            for side in (0, 1):
                assert torch.equal(eng_a.covariances(side), eng_b.covariances(side))
                assert torch.equal(eng_a.knn(side)[0], eng_b.knn(side)[0])
    # a batch of small pairs with different sizes per side
    offs_t = np.array([0, 300, 300 + 217, 300 + 217 + 360], dtype=np.int64)
    offs_s = np.array([0, 280, 280 + 360, 280 + 360 + 90], dtype=np.int64)
    rng = np.random.default_rng(0)
    tb = torch.as_tensor(rng.uniform(0, 500, (int(offs_t[-1]), 2)), device="cuda")
    sb = torch.as_tensor(rng.uniform(0, 500, (int(offs_s[-1]), 2)), device="cuda")
    eng_a.set_target(tb, offs_t); eng_a.set_source(sb, offs_s)
    eng_b.set_pair(tb, sb, offs_t, offs_s)
    for side in (0, 1):
        assert torch.equal(eng_a.covariances(side), eng_b.covariances(side))
    assert torch.equal(eng_a.register(history=False).T, eng_b.register(history=False).T)
    # larger clouds (multi-kernel grid build, sort scratch per side): side by side, and the sequential branch of the
    # same entry point (GICP_PAIR_OVERLAP_MAX=0)
    s3, t3, _ = synthetic.patches3d_pair(n=20000, n_patches=8, cube=40.0, patch=30.0, seed=4)
    e3a, e3b = GicpEngine(3, "f32"), GicpEngine(3, "f32")
    for e in (e3a, e3b):
        e.set_params(k=20, max_distance_nearest_neighbors=5.0, max_distance_correspondence=2.0)
    s3 = torch.as_tensor(s3, dtype=torch.float32, device="cuda"); t3 = torch.as_tensor(t3, dtype=torch.float32, device="cuda")
    e3a.set_target(t3); e3a.set_source(s3)
    Ta = e3a.register(history=False).T
    for rep in range(3):
        e3b.set_pair(t3, s3)
        for side in (0, 1):
            assert torch.equal(e3a.covariances(side), e3b.covariances(side))
            assert torch.equal(e3a.knn(side)[0], e3b.knn(side)[0])
        assert torch.equal(Ta, e3b.register(history=False).T)
    saved = os.environ.get("GICP_PAIR_OVERLAP_MAX")
    os.environ["GICP_PAIR_OVERLAP_MAX"] = "0"
    try:
        e3b.set_pair(t3, s3)
        assert torch.equal(Ta, e3b.register(history=False).T)
    finally:
        if saved is None:
            os.environ.pop("GICP_PAIR_OVERLAP_MAX", None)
        else:
            os.environ["GICP_PAIR_OVERLAP_MAX"] = saved
