"""Stand-in for robot-visualization.py's process layout (lines 151-166, 195-200): the parent imports
the module and only calls apply_transformation; a FORKED worker makes the first gicp() call.
Run as a script in a fresh interpreter; prints 'FORK-OK' on success."""
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gicp as shim  # noqa: E402  (parent: import only, nothing CUDA-related may run)


def worker(q_in, q_out):
    src, tgt = q_in.get()
    T, _, c_src, c_tgt, _, _, _ = shim.gicp(src, tgt, max_distance_nearest_neighbors=200, tolerance=1)
    q_out.put((T, c_src, c_tgt))


def main():
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "config2_rays90_pair3.npz")))
    assert np.allclose(shim.apply_transformation(g["src"], np.eye(3)), g["src"])
    ctx = mp.get_context("fork")
    q_in, q_out = ctx.Queue(), ctx.Queue()
    p = ctx.Process(target=worker, args=(q_in, q_out))
    p.start()
    q_in.put(([tuple(x) for x in g["src"]], [tuple(x) for x in g["tgt"]]))
    T, c_src, c_tgt = q_out.get(timeout=180)
    p.join(timeout=30)
    assert p.exitcode == 0
    assert T.shape == (3, 3) and np.isfinite(T).all()
    assert np.abs(c_tgt - g["tgt_cov"]).max() < 1e-9 and np.abs(c_src - g["src_cov0"]).max() < 1e-9
    assert np.isfinite([-T[0, 2], -T[1, 2], -np.arctan2(T[1, 0], T[0, 0])]).all()
    print("FORK-OK")


if __name__ == "__main__":
    main()
