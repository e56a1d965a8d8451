set -x
timeout 400 python -m pytest tests/test_gpu_sharded.py -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_sharded_n8.log
for N in ${NS:-8 4}; do
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"
timeout 300 $R bench.py --gpus $N --workload config5 --steps 2 --warmup 3 > gpurun_out/bench_c5_n$N.json 2> gpurun_out/bench_c5_n$N.err; tail -2 gpurun_out/bench_c5_n$N.err
timeout 300 $R bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; tail -2 gpurun_out/bench_c4_n$N.err
done
python - <<'PY'
import json
import os
for N in [int(x) for x in os.environ.get("NS", "8 4").split()]:
    try:
        d=json.load(open(f"gpurun_out/bench_c5_n{N}.json")); print("c5",N,d["ms_per_step"],d["value"],d["e2e"]["value"],{k:[round(x,2) for x in v] for k,v in d["roofline"]["stage_ms_per_rank"].items()})
    except Exception as e: print("c5",N,e)
    try:
        d=json.load(open(f"gpurun_out/bench_c4_n{N}.json")); print("c4",N,d["ms_per_step"],d["value"],d["pairs_per_sec"],d["scaling"],d["e2e"]["pairs_per_sec"],d["hbm_used_gb_max_rank"])
    except Exception as e: print("c4",N,e)
PY
