"""Latency / throughput of the other BASELINE.json configs (parity-test cases, not the bench line):
config 1 (visualization.py pair), config 2 (robot scan sequence), config 3 (100k 3-D pair),
config 5 (16M 3-D pair; sharded over the ranks when launched with torchrun).
Prints one JSON object per config.  Needs a GPU."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from generalized_icp_b200 import compat, synthetic  # noqa: E402
from generalized_icp_b200.engine import GicpEngine  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tests"))
import demo_inputs  # noqa: E402  (seeded restatements of the demos' input recipes)

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
dev = torch.device("cuda", torch.cuda.current_device())
which = sys.argv[1:] or ["1", "2", "3"]


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


if "1" in which and rank == 0:
    ts, its = [], []
    for seed in range(8):
        s, t = demo_inputs.config1_pair(seed)
        dt, r = timed(lambda: compat.gicp_extended(s, t, full_history=False), 5)
        ts.append(dt)
        its.append(r["n_outer"])
    print(json.dumps({"config": 1, "what": "visualization.py pair, 90 x 87 points, defaults, host arrays in / host results out",
                      "ms_per_pair_median": 1e3 * float(np.median(ts)), "outer_iterations": its}))

if "2" in which and rank == 0:
    for rays in (90, 360):
        scans, _ = demo_inputs.lidar_sequence(seed=1, num_rays=rays, n_scans=30)
        pairs = [(np.asarray(scans[i]), np.asarray(scans[i + 1])) for i in range(len(scans) - 1)]

        def run():
            return [compat.gicp_extended(a, b, max_distance_nearest_neighbors=200, tolerance=1, full_history=False)["n_outer"]
                    for a, b in pairs]
        dt, its = timed(run, 3)
        print(json.dumps({"config": 2, "what": f"robot scan sequence, {rays} rays, {len(pairs)} consecutive pairs, r_knn 200, tol 1",
                          "ms_per_pair": 1e3 * dt / len(pairs), "pairs_per_sec": len(pairs) / dt,
                          "mean_outer_iterations": float(np.mean(its))}))

if "2" in which and rank == 0:
    # the call the robot demo makes (robot-visualization.py:157-162): the full 7-tuple with every iteration's
    # covariances and highest-weight correspondences
    import gicp as shim
    import contextlib
    import io
    scans, _ = demo_inputs.lidar_sequence(seed=1, num_rays=360, n_scans=30)
    pairs = [([tuple(p) for p in scans[i]], [tuple(p) for p in scans[i + 1]]) for i in range(len(scans) - 1)]

    def run7():
        with contextlib.redirect_stdout(io.StringIO()):
            return [len(shim.gicp(a, b, max_distance_nearest_neighbors=200, tolerance=1)[1]) for a, b in pairs]
    dt, _ = timed(run7, 3)
    print(json.dumps({"config": 2, "what": "robot scan sequence, 360 rays: gicp() with the reference's full 7-tuple (lists of tuples in)",
                      "ms_per_pair": 1e3 * dt / len(pairs)}))

if "3" in which and rank == 0:
    src, tgt, T = synthetic.patches3d_pair(**synthetic.CONFIG3, seed=0)
    eng = GicpEngine(3, "f32")
    eng.set_params(**synthetic.CONFIG3_PARAMS)
    s_d, t_d = torch.as_tensor(src, device=dev), torch.as_tensor(tgt, device=dev)

    def run():
        eng.set_pair(t_d, s_d)     # = set_target + set_source; the two set-ups overlap on the device
        return eng.register(history=False)
    dt, r = timed(run, 5)
    n_it = int(r.n_outer[0])
    Te = r.T[0].cpu().numpy()
    ang = float(np.arccos(np.clip((np.trace(Te[:3, :3].T @ T[:3, :3]) - 1) / 2, -1, 1)))
    print(json.dumps({"config": 3, "what": "single 3-D pair, 100k points per side, k 20, device-resident inputs",
                      "ms_per_pair": 1e3 * dt, "outer_iterations": n_it, "correspondences_per_sec": 100000 * n_it / dt,
                      "rot_err_rad": ang, "trans_err_m": float(np.linalg.norm(Te[:3, 3] - T[:3, 3]))}))

if "5" in which:
    import torch.distributed as dist
    n = int(os.environ.get("CONFIG5_N", synthetic.CONFIG5["n"]))
    cfg = dict(synthetic.CONFIG5, n=n)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    src, tgt, _, Tt = synthetic.patches3d_batch_device(1, seed=5, device=dev, chunk=1, **{k: v for k, v in cfg.items()})
    eng = GicpEngine(3, "f32", device=dev.index)
    if world > 1:
        uid = [GicpEngine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(world, rank, uid[0])
    eng.set_params(**synthetic.CONFIG5_PARAMS)

    def run():
        eng.set_target(tgt)
        eng.set_source(src)
        return eng.register(history=False)
    run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = run()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    n_it = int(r.n_outer[0])
    Te, Tg = r.T[0].cpu().numpy(), Tt[0].cpu().numpy()
    ang = float(np.arccos(np.clip((np.trace(Te[:3, :3].T @ Tg[:3, :3]) - 1) / 2, -1, 1)))
    if rank == 0:
        print(json.dumps({"config": 5, "what": f"single 3-D pair, {n} points per side, source sharded over {world} rank(s), "
                          "NCCL all-gather once + all-reduce of 80 f64 per outer iteration",
                          "n_gpus": world, "ms_per_pair": float(ms), "outer_iterations": n_it,
                          "correspondences_per_sec": n * n_it / (float(ms) * 1e-3), "rot_err_rad": ang,
                          "trans_err_m": float(np.linalg.norm(Te[:3, 3] - Tg[:3, 3]))}))
    if world > 1:
        eng.comm_destroy()
        dist.destroy_process_group()
