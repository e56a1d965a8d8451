"""End-to-end outcomes of the UNMODIFIED reference on the BASELINE configs 1-2, for the match-rate report.

Run in the dev container only (needs /root/reference):

    python tests/golden/make_match_rate.py

Cases: config 1 (visualization.py's pair) seeds 0-39 with the defaults of gicp.py:78; config 2, a 30-scan
scripted drive at 90 and at 360 rays, consecutive pairs with the robot demo's parameters
(robot-visualization.py:160-161).  Inputs are NOT stored - ``tests/demo_inputs.py`` regenerates them
bit for bit from the seed; only the reference's outcome is: iteration count, final transform, whether
every fmin_cg call returned warnflag 0, and the oracle's outcome with a converged inner solve on the same
input (so the CPU suite can tell a regression of the engine from the reference's own chaos).
Nothing from the reference is copied; the file is stamped with the numpy / scipy versions."""
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np
import scipy

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference/python-implementation")
import gicp as ref  # noqa: E402  (the reference module)

sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import demo_inputs  # noqa: E402
from oracle import gicp_oracle as O  # noqa: E402


def cases():
    for seed in range(40):
        s, t = demo_inputs.config1_pair(seed)
        yield ("config1", seed, 0, 0), s, t, {}
    for rays in (90, 360):
        scans, _ = demo_inputs.lidar_sequence(seed=1, num_rays=rays, n_scans=30)
        for i in range(len(scans) - 1):
            yield (("config2", 1, rays, i), np.asarray(scans[i]), np.asarray(scans[i + 1]),
                   dict(max_distance_nearest_neighbors=200, tolerance=1))


def run_reference(src, tgt, kw):
    flags = []
    real = ref.fmin_cg

    def spy(f, x0, fprime, **k):
        out = real(f=f, x0=x0, fprime=fprime, **k)
        flags.append(int(out[4]))
        return out

    ref.fmin_cg = spy
    try:
        with redirect_stdout(io.StringIO()):
            out = ref.gicp(src, tgt, **kw)
    finally:
        ref.fmin_cg = real
    return out[0], len(out[6]), flags


def main():
    kind, seed, rays, pair, n_ref, T_ref, clean, n_newton, T_newton = [], [], [], [], [], [], [], [], []
    for (kd, sd, ry, pr), s, t, kw in cases():
        T, n, flags = run_reference(s, t, kw)
        o = O.gicp_oracle(s, t, inner="newton", recompute_src_cov=True, record=False, **kw)
        kind.append(1 if kd == "config1" else 2)
        seed.append(sd); rays.append(ry); pair.append(pr)
        n_ref.append(n); T_ref.append(T); clean.append(all(f == 0 for f in flags))
        n_newton.append(o["n_outer"]); T_newton.append(o["T"])
        print(kd, sd, ry, pr, "reference iterations", n, "warnflags clean", clean[-1], "converged-inner iterations", o["n_outer"])
    np.savez_compressed(os.path.join(HERE, "match_rate_reference.npz"), kind=np.array(kind), seed=np.array(seed),
                        rays=np.array(rays), pair=np.array(pair), n_ref=np.array(n_ref), T_ref=np.stack(T_ref),
                        clean=np.array(clean), n_newton=np.array(n_newton), T_newton=np.stack(T_newton),
                        versions=f"numpy {np.__version__} scipy {scipy.__version__}")


if __name__ == "__main__":
    main()
