#!/bin/bash
# Round-end evidence on ONE B200 (run under gpurun): tests, the bench line, the reference arm, the ncu launch list
# (the same bench command at 512 pairs), one ncu --set full capture of a 64-pair step, the other configs.
set -x
mkdir -p gpurun_out
KREGEX='regex:knn_|correspond|accumulate|solve_|presum|init_state|bbox_|grid_meta|morton_lut|cell_key|gather_sorted|regather|fill_cov|DeviceRadixSort|DeviceScan|small_grid|register_loop'
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r02.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r02.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo "bench rc=$?"
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_r02.json 2> gpurun_out/bench_ref_r02.err; echo "ref rc=$?"
python bench.py --pairs 512 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name "$KREGEX" -c 4000 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --pairs 512 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
python scripts/profile_step.py 64 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name 'regex:knn_hist|correspond_kernel|accumulate_kernel|solve_kernel' -c 11 \
    -o gpurun_out/prof_r2_final -f python scripts/profile_step.py 64 > gpurun_out/ncu.log 2>&1
python scripts/bench_configs.py 1 2 3 > gpurun_out/configs_r02.jsonl 2> gpurun_out/configs_r02.err
(python scripts/odometry_eval.py 360 36; python scripts/odometry_eval.py 90 36) > gpurun_out/odometry_r02.json 2> gpurun_out/odometry_r02.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["pairs_per_sec"], d["e2e"]["value"], d["e2e"]["pairs_per_sec"],
      {k: round(v["ms_per_step"], 2) for k, v in d["roofline"]["kernels"].items()},
      {k: round(v["frac"], 4) for k, v in d["roofline"]["kernels"].items()}, d["cpu_baseline"]["value"], d["hbm_used_gb_max_rank"], d["clocks"])
PY
cat gpurun_out/configs_r02.jsonl | cut -c1-250
