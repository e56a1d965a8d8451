"""One registration step of the bench workload on a small batch - the target of ncu captures.
    python scripts/profile_step.py [pairs] [points] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from generalized_icp_b200 import synthetic  # noqa: E402
from generalized_icp_b200.engine import GicpEngine  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
points = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cfg = {k: v for k, v in synthetic.CONFIG4.items() if k != "n"}
src, tgt, off, _ = synthetic.patches3d_batch_device(pairs, n=points, seed=0, device="cuda", **cfg)
off = off.cpu().numpy()
eng = GicpEngine(3, "f32")
eng.set_params(**synthetic.CONFIG4_PARAMS, knn_cell=float(os.environ.get("KNN_CELL", 0)), nn_cell=float(os.environ.get("NN_CELL", 0)))
for _ in range(steps):
    eng.set_target(tgt, off)
    eng.set_source(src, off)
    r = eng.register(history=False)
torch.cuda.synchronize()
print("n_outer mean", float(r.n_outer.double().mean()), "launches", eng.launch_count)
