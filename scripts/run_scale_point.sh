#!/bin/bash
# one rank count of the scaling tables: NS="4" bash scripts/run_scale_point.sh   (config 5 + config 4 bench lines)
for N in ${NS:-4}; do
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"
[ "$N" = "1" ] && R="python"
timeout 300 $R bench.py --gpus $N --workload config5 --steps 2 --warmup 3 > gpurun_out/bench_c5_n$N.json 2> gpurun_out/bench_c5_n$N.err; tail -2 gpurun_out/bench_c5_n$N.err
[ "$N" != "1" ] && { timeout 300 $R bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; tail -2 gpurun_out/bench_c4_n$N.err; }
done
python - <<'PY'
import json, os
for N in [int(x) for x in os.environ.get("NS", "4").split()]:
    for c in ("c5", "c4"):
        try:
            d = json.loads(open(f"gpurun_out/bench_{c}_n{N}.json").read().strip().splitlines()[-1])
            print(c, N, round(d["ms_per_step"], 2), round(d.get("pairs_per_sec", 0)), round(d["e2e"].get("pairs_per_sec", d["e2e"]["value"])),
                  {k: round(v["ms_per_step"], 2) for k, v in d["roofline"]["kernels"].items()})
        except Exception as e:
            print(c, N, "-", e)
PY
