#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== latency path"; python scripts/latency_breakdown.py 360; python scripts/latency_breakdown.py 90
python scripts/bench_configs.py 1 2 3 > gpurun_out/configs_r2b.jsonl 2>gpurun_out/configs_r2b.err; cat gpurun_out/configs_r2b.jsonl | cut -c1-260
echo "== bench 1024 pairs"; python bench.py --pairs 1024 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1024.json 2> gpurun_out/bench_1024.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_1024.json').read().strip().splitlines()[-1])
print('ms', d['ms_per_step'], 'pairs/s', d['pairs_per_sec'], 'e2e pairs/s', d['e2e']['pairs_per_sec'], {k:round(v['ms_per_step'],2) for k,v in d['roofline']['kernels'].items()})
PY
