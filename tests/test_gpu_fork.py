"""The reference's robot demo imports the module in the UI process and first CALLS gicp() in a
forked worker, with lists of tuples in and the result pickled back through a Queue
(robot-visualization.py:151-166,195-200).  CUDA must therefore initialise lazily, inside the call."""
import multiprocessing as mp

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(q_in, q_out):
    import gicp as shim                      # inherited, already imported by the parent
    src, tgt = q_in.get()
    T, _, c_src, c_tgt, _, _, _ = shim.gicp(src, tgt, max_distance_nearest_neighbors=200, tolerance=1)
    q_out.put((T, c_src, c_tgt))


def test_first_call_in_forked_worker(golden):
    import gicp as shim                      # parent: import only (no CUDA)
    g = golden("config2_rays90_pair3")
    T_ui = np.eye(3)
    assert np.allclose(shim.apply_transformation(g["src"], T_ui), g["src"])   # UI-side call is pure numpy
    ctx = mp.get_context("fork")
    q_in, q_out = ctx.Queue(), ctx.Queue()
    p = ctx.Process(target=_worker, args=(q_in, q_out))
    p.start()
    q_in.put(([tuple(x) for x in g["src"]], [tuple(x) for x in g["tgt"]]))
    T, c_src, c_tgt = q_out.get(timeout=120)
    p.join(timeout=30)
    assert p.exitcode == 0
    assert T.shape == (3, 3) and np.isfinite(T).all()
    assert np.abs(c_tgt - g["tgt_cov"]).max() < 1e-9 and np.abs(c_src - g["src_cov0"]).max() < 1e-9
    # pose integration as the demo does it (robot-visualization.py:258-260)
    assert np.isfinite([-T[0, 2], -T[1, 2], -np.arctan2(T[1, 0], T[0, 0])]).all()
