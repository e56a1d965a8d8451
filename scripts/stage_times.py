"""Per-stage milliseconds of one registration step of the bench workload (profiled run, CUDA events per stage).
    python scripts/stage_times.py [pairs] [repeats]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from generalized_icp_b200 import synthetic  # noqa: E402
from generalized_icp_b200.engine import GicpEngine  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = {k: v for k, v in synthetic.CONFIG4.items() if k != "n"}
src, tgt, off, _ = synthetic.patches3d_batch_device(pairs, n=32768, seed=0, device="cuda", **cfg)
off = off.cpu().numpy()
eng = GicpEngine(3, "f32")
eng.set_params(**synthetic.CONFIG4_PARAMS)


def step():
    eng.set_target(tgt, off)
    eng.set_source(src, off)
    return eng.register(history=False)


step()
eng.profile(True)
eng.profile_read()
for _ in range(reps):
    r = step()
torch.cuda.synchronize()
p = eng.profile_read()
tot = 0.0
for s, (ms, n) in p.items():
    print(f"{s:12s} {ms / reps:9.2f} ms  ({n // reps} sections)")
    tot += ms / reps
print(f"{'sum':12s} {tot:9.2f} ms per {pairs} pairs; n_outer mean {float(r.n_outer.double().mean()):.2f}")
