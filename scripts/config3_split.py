"""Where one 100k + 100k pair (BASELINE configs[2]) spends its time: wall split with synchronisations, per-stage device
times, per-side k-NN, and gicpSetPair (set-ups side by side) against the two separate calls.
    python scripts/config3_split.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from generalized_icp_b200 import synthetic
from generalized_icp_b200.engine import GicpEngine
src, tgt, T = synthetic.patches3d_pair(**synthetic.CONFIG3, seed=0)
eng = GicpEngine(3, "f32"); eng.set_params(**synthetic.CONFIG3_PARAMS)
s_d, t_d = torch.as_tensor(src, device="cuda"), torch.as_tensor(tgt, device="cuda")
def run():
    eng.set_target(t_d); eng.set_source(s_d); return eng.register(history=False)
for _ in range(3): run()
torch.cuda.synchronize()
# wall split with syncs
acc = {}
for _ in range(5):
    t0 = time.perf_counter(); eng.set_target(t_d); torch.cuda.synchronize(); t1 = time.perf_counter()
    eng.set_source(s_d); torch.cuda.synchronize(); t2 = time.perf_counter()
    r = eng.register(history=False); torch.cuda.synchronize(); t3 = time.perf_counter()
    for k, v in (("set_target", t1 - t0), ("set_source", t2 - t1), ("register", t3 - t2)): acc[k] = acc.get(k, 0) + v / 5
print({k: round(1e3 * v, 3) for k, v in acc.items()}, "n_outer", int(r.n_outer[0]))
eng.profile(True); eng.profile_read()
for _ in range(3): run()
torch.cuda.synchronize()
p = eng.profile_read()
print({k: (round(v[0] / 3, 3), v[1] // 3) for k, v in p.items()})
for name, fn in (("target", lambda: eng.set_target(t_d)), ("source", lambda: eng.set_source(s_d)), ("target again as source", lambda: eng.set_source(t_d)), ("source as target", lambda: eng.set_target(s_d))):
    eng.profile_read()
    for _ in range(3): fn()
    torch.cuda.synchronize()
    p = eng.profile_read()
    print(name, {k: round(v[0] / 3, 3) for k, v in p.items() if v[1]}, "launches per call", None)
l0 = eng.launch_count; eng.set_target(t_d); l1 = eng.launch_count; eng.set_source(s_d); l2 = eng.launch_count
print("launches target", l1 - l0, "source", l2 - l1)

eng.profile(False)   # per-stage timing forces the sequential branch of gicpSetPair


def run_pair():
    eng.set_pair(t_d, s_d); return eng.register(history=False)
for fn, name in ((run, "set_target + set_source + register"), (run_pair, "set_pair + register")):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): r = fn()
    torch.cuda.synchronize()
    print(f"{name}: {1e2 * (time.perf_counter() - t0):.3f} ms per pair, n_outer {int(r.n_outer[0])}")
ra = run().T.clone(); rb = run_pair().T
print("identical T:", bool(torch.equal(ra, rb)))
