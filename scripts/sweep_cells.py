"""Sweep grid cell sizes on a small batch of the bench workload; prints per-stage ms."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from generalized_icp_b200 import synthetic
from generalized_icp_b200.engine import GicpEngine

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 128
cfg = {k: v for k, v in synthetic.CONFIG4.items() if k != "n"}
src, tgt, off, _ = synthetic.patches3d_batch_device(pairs, n=32768, seed=0, device="cuda", **cfg)
off = off.cpu().numpy()
eng = GicpEngine(3, "f32")
combos = [(0.0, n) for n in (0.0, 1.5)]
for knn_cell, nn_cell in combos:
    eng.set_params(**synthetic.CONFIG4_PARAMS, knn_cell=knn_cell, nn_cell=nn_cell)
    for rep in range(2):
        eng.profile(True)
        eng.set_target(tgt, off); eng.set_source(src, off); r = eng.register(history=False)
        p = eng.profile_read()
    eng.profile(False)
    print(f"knn_cell={knn_cell} nn_cell={nn_cell} n_outer={float(r.n_outer.double().mean()):.2f} " +
          " ".join(f"{k}={v[0]:.1f}ms" for k, v in p.items()), flush=True)
