"""Does registering two halves of a batch on two streams (two handles) overlap the issue-bound k-NN of one half
with the latency/bandwidth-bound outer loop of the other?   python scripts/overlap_test.py [pairs] [chunks]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from generalized_icp_b200 import synthetic  # noqa: E402
from generalized_icp_b200.engine import GicpEngine  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n_chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 4
N = 32768
cfg = {k: v for k, v in synthetic.CONFIG4.items() if k != "n"}
src, tgt, off, _ = synthetic.patches3d_batch_device(pairs, n=N, seed=0, device="cuda", **cfg)
off_h = off.cpu().numpy()


def timed(fn, reps=3):
    fn(); fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, out


eng = GicpEngine(3, "f32")
eng.set_params(**synthetic.CONFIG4_PARAMS)


def plain():
    eng.set_target(tgt, off_h)
    eng.set_source(src, off_h)
    return eng.register(history=False).n_outer


ms0, n0 = timed(plain)
print(f"one handle, one stream: {ms0:.1f} ms per {pairs} pairs", flush=True)

per = pairs // n_chunks
engs = [GicpEngine(3, "f32"), GicpEngine(3, "f32")]
for e in engs:
    e.set_params(**synthetic.CONFIG4_PARAMS)
for prio in ((0, 0), (-1, 0)):
    streams = [torch.cuda.Stream(priority=prio[0]), torch.cuda.Stream(priority=prio[1])]
    loc = off_h[:per + 1]

    def overlapped():
        outs = []
        main = torch.cuda.current_stream()
        for s in streams:
            s.wait_stream(main)

        def setc(c):
            e, s = engs[c % 2], streams[c % 2]
            with torch.cuda.stream(s):
                e.set_target(tgt[c * per * N:(c + 1) * per * N], loc)
                e.set_source(src[c * per * N:(c + 1) * per * N], loc)

        setc(0)
        for c in range(n_chunks):
            if c + 1 < n_chunks:
                setc(c + 1)                      # queued on the other stream before this chunk's loop blocks the host
            with torch.cuda.stream(streams[c % 2]):
                outs.append(engs[c % 2].register(history=False).n_outer)
        for s in streams:
            main.wait_stream(s)
        return torch.cat(outs)

    ms1, n1 = timed(overlapped)
    print(f"two handles, two streams (priorities {prio}), {n_chunks} chunks: {ms1:.1f} ms  ({ms0 / ms1:.3f}x), same iteration counts: {bool(torch.equal(n0, n1))}", flush=True)
