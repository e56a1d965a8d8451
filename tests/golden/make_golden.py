"""Generate the golden fixtures by running the UNMODIFIED reference.

Run in the dev container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case it imports ``/root/reference/python-implementation/gicp.py`` by
path, runs ``gicp.gicp`` on seeded inputs and records
  * the returned 7-tuple (gicp.py:174),
  * what the inner optimiser saw on every outer iteration, captured by wrapping
    the module's ``fmin_cg`` name (the closure cells of the two lambdas at
    gicp.py:148-149 hold corresponding_target_points and weight_matrices),
  * the k-NN index lists of both clouds, obtained with the same scipy call the
    reference makes (gicp.py:24), and the 1-NN index per outer iteration
    (gicp.py:132).
Nothing from the reference is copied; the outputs are stamped with the numpy /
scipy versions that produced them.
"""
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np
import scipy
from scipy.spatial import KDTree

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference/python-implementation")
import gicp as ref  # noqa: E402  (the reference module)

sys.path.insert(0, os.path.join(ROOT, "tests"))
import demo_inputs  # noqa: E402


def cells(fn):
    return dict(zip(fn.__code__.co_freevars, [c.cell_contents for c in fn.__closure__]))


def run_case(src, tgt, **kw):
    src = np.asarray(src, dtype=np.float64)
    tgt = np.asarray(tgt, dtype=np.float64)
    log = []
    real = ref.fmin_cg

    def spy(f, x0, fprime, **k):
        out = real(f=f, x0=x0, fprime=fprime, **k)
        c = cells(f)
        x = out[0]
        log.append(dict(x0=np.array(x0), xopt=np.array(x), fopt=float(out[1]), fcalls=int(out[2]),
                        warnflag=int(out[4]), q=c["corresponding_target_points"].copy(),
                        W=c["weight_matrices"].copy(), loss_at_xopt=float(f(x)), grad_at_xopt=np.array(fprime(x)),
                        loss_at_x0=float(f(np.array(x0))), grad_at_x0=np.array(fprime(np.array(x0)))))
        return out

    ref.fmin_cg = spy
    buf = io.StringIO()
    try:
        with redirect_stdout(buf):
            T, all_T, c_src0, c_tgt, hw_s, hw_t, all_c = ref.gicp(src, tgt, **kw)
    finally:
        ref.fmin_cg = real
    r_knn = kw.get("max_distance_nearest_neighbors", 50)
    d_max = kw.get("max_distance_correspondence", 150)

    def knn(points):
        tree = KDTree(points)
        return np.stack([tree.query(points[i], k=6, distance_upper_bound=r_knn)[1] for i in range(len(points))])

    tree = KDTree(tgt)
    nn_idx, nn_dist = [], []
    for k_it in range(len(log)):
        moved = ref.apply_transformation(src, all_T[k_it] if k_it < len(all_T) else T)
        d, j = tree.query(moved)
        nn_idx.append(np.where(d > d_max, -1, j))
        nn_dist.append(d)
    out = dict(src=src, tgt=tgt, T=T, all_T=np.stack(all_T), src_cov0=c_src0, tgt_cov=c_tgt,
               all_src_cov=np.stack(all_c), n_hw=len(hw_s), stdout=buf.getvalue(),
               src_knn=knn(src), tgt_knn=knn(tgt), nn_idx=np.stack(nn_idx), nn_dist=np.stack(nn_dist),
               r_knn=float(r_knn), d_max=float(d_max), tolerance=float(kw.get("tolerance", 1e-6)),
               max_iterations=int(kw.get("max_iterations", 100)),
               versions=f"numpy {np.__version__} scipy {scipy.__version__}")
    for i, (a, b) in enumerate(zip(hw_s, hw_t)):
        out[f"hw_src_{i}"] = a
        out[f"hw_tgt_{i}"] = b
    for key in ("x0", "xopt", "fopt", "fcalls", "warnflag", "q", "W", "loss_at_xopt", "grad_at_xopt",
                "loss_at_x0", "grad_at_x0"):
        out["it_" + key] = np.stack([np.asarray(e[key]) for e in log])
    return out


def main():
    made = []
    for seed in range(6):
        s, t = demo_inputs.config1_pair(seed)
        made.append((f"config1_seed{seed}", run_case(s, t)))
    for rays, n_scans in ((90, 11), (360, 7)):
        scans, _ = demo_inputs.lidar_sequence(seed=1, num_rays=rays, n_scans=n_scans)
        for i in range(len(scans) - 1):
            # source = previous scan, target = current (robot-visualization.py:250-251)
            made.append((f"config2_rays{rays}_pair{i}",
                         run_case(scans[i], scans[i + 1], max_distance_nearest_neighbors=200, tolerance=1)))
    for name, data in made:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
        print(name, "outer iterations:", len(data["it_fopt"]), "warnflags:", data["it_warnflag"].tolist(),
              data["stdout"].strip())


if __name__ == "__main__":
    main()
