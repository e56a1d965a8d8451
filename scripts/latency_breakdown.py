"""Where a small registration spends its wall time (config 2: one pair of 360-beam scans).
    python scripts/latency_breakdown.py            (GICP_FUSED_LOOP=0 for the multi-launch loop)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import demo_inputs  # noqa: E402
from generalized_icp_b200 import compat  # noqa: E402
from generalized_icp_b200.engine import GicpEngine  # noqa: E402

rays = int(sys.argv[1]) if len(sys.argv) > 1 else 360
scans, _ = demo_inputs.lidar_sequence(seed=1, num_rays=rays, n_scans=30)
pairs = [(np.asarray(scans[i], dtype=np.float64), np.asarray(scans[i + 1], dtype=np.float64)) for i in range(len(scans) - 1)]
eng = GicpEngine(2, "f64")
eng.set_params(k=6, max_distance_nearest_neighbors=200.0, max_distance_correspondence=150.0, tolerance=1.0,
               inner_max_iterations=int(os.environ.get("INNER_MAX", 50)))
dev = eng.device
acc = {}


def tick(name, t0):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    acc[name] = acc.get(name, 0.0) + (t1 - t0)
    return t1


for rep in range(4):
    if rep == 1:
        acc.clear()
    for a, b in pairs:
        t = time.perf_counter()
        s_dev = torch.as_tensor(a, device=dev)
        t_dev = torch.as_tensor(b, device=dev)
        t = tick("h2d", t)
        l0 = eng.launch_count
        eng.set_target(t_dev)
        t = tick("set_target", t)
        eng.set_source(s_dev)
        t = tick("set_source", t)
        l1 = eng.launch_count
        r = eng.register(history=False)
        t = tick("register", t)
        l2 = eng.launch_count
        T = r.T.cpu()
        n = int(r.n_outer[0])
        t = tick("readback", t)
n = 3 * len(pairs)
print(f"rays {rays}: " + "  ".join(f"{k} {1e6 * v / n:.0f} us" for k, v in acc.items()) +
      f"  | total {1e6 * sum(acc.values()) / n:.0f} us per pair; launches set {l1 - l0} register {l2 - l1}")
t0 = time.perf_counter()
for a, b in pairs:
    compat.gicp_extended(a, b, max_distance_nearest_neighbors=200, tolerance=1, full_history=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    for a, b in pairs:
        compat.gicp_extended(a, b, max_distance_nearest_neighbors=200, tolerance=1, full_history=False)
torch.cuda.synchronize()
print(f"compat.gicp_extended(full_history=False): {1e6 * (time.perf_counter() - t0) / n:.0f} us per pair")
