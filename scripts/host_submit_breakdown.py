"""Host-side submission time of every step of GicpEngine.register_pair_host (no synchronisation between the steps):
where the host, not the device, bounds a small registration.    python scripts/host_submit_breakdown.py [rays]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import demo_inputs  # noqa: E402
from generalized_icp_b200 import _lib  # noqa: E402
from generalized_icp_b200.engine import GicpEngine, SOURCE, TARGET  # noqa: E402

rays = int(sys.argv[1]) if len(sys.argv) > 1 else 90
scans, _ = demo_inputs.lidar_sequence(seed=1, num_rays=rays, n_scans=30)
pairs = [(np.asarray(scans[i], dtype=np.float64), np.asarray(scans[i + 1], dtype=np.float64)) for i in range(len(scans) - 1)]
eng = GicpEngine(2, "f64")
eng.set_params(k=6, max_distance_nearest_neighbors=200.0, max_distance_correspondence=150.0, tolerance=1.0)
acc = {}
for a, b in pairs:
    eng.register_pair_host(a, b)
torch.cuda.synchronize()


def stamp(name, t0):
    t1 = time.perf_counter()
    acc[name] = acc.get(name, 0.0) + (t1 - t0)
    return t1


d, d1 = 2, 3
mi = int(eng.params.max_iterations)
vp = C.c_void_p
for rep in range(3):
    for src, tgt in pairs:
        t = time.perf_counter()
        n_s, n_t = src.shape[0], tgt.shape[0]
        pad_s = (n_s + 3) & ~3
        rows = pad_s + n_t
        hv = eng._pp_host.numpy()
        hv[:n_s] = src
        hv[pad_s:rows] = tgt
        t = stamp("stage into pinned", t)
        eng._pp_dev[:rows].copy_(eng._pp_host[:rows], non_blocking=True)
        t = stamp("h2d submit", t)
        st = eng._stream()
        base, row_bytes = eng._pp_dev.data_ptr(), d * eng._pp_dev.element_size()
        _lib.check(eng.lib.gicpSetPair(eng._h, vp(base + pad_s * row_bytes), (C.c_int64 * 2)(0, n_t), vp(base),
                                       (C.c_int64 * 2)(0, n_s), 1, st))
        t = stamp("set_pair", t)
        sizes = [d1 * d1, mi, (mi + 1) * d1 * d1, n_s * d * d, n_t * d * d]
        offs = [0]
        for z in sizes:
            offs.append(offs[-1] + z)
        nd = offs[-1]
        tot = nd + (2 + mi + 1) // 2
        dp = eng._po_dev.data_ptr()
        ip = dp + 8 * nd
        _lib.check(eng.lib.gicpRegister(eng._h, None, vp(dp), vp(ip), vp(ip + 4), vp(dp + 8 * offs[1]),
                                        vp(dp + 8 * offs[2]), vp(ip + 8), st))
        t = stamp("gicpRegister", t)
        _lib.check(eng.lib.gicpCovariances(eng._h, SOURCE, vp(dp + 8 * offs[3]), st))
        _lib.check(eng.lib.gicpCovariances(eng._h, TARGET, vp(dp + 8 * offs[4]), st))
        t = stamp("gicpCovariances x2", t)
        eng._po_host[:tot].copy_(eng._po_dev[:tot], non_blocking=True)
        torch.cuda.current_stream(eng.device).synchronize()
        t = stamp("d2h (waits for the device)", t)
        hd = eng._po_host.numpy()[:tot].copy()
        t = stamp("copy out of the pinned mirror", t)
n = 3 * len(pairs)
print(f"rays {rays}, host microseconds per pair: " + "; ".join(f"{k} {1e6 * v / n:.1f}" for k, v in acc.items()) +
      f"; total {1e6 * sum(acc.values()) / n:.1f}")
