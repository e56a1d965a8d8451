#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== latency path"; python scripts/latency_breakdown.py 360; python scripts/latency_breakdown.py 90
echo "== multi-launch"; GICP_FUSED_LOOP=0 GICP_SMALL_GRID=0 python scripts/latency_breakdown.py 360
python scripts/bench_configs.py 1 2 3 > gpurun_out/configs_r2b.jsonl 2>gpurun_out/configs_r2b.err; cat gpurun_out/configs_r2b.jsonl | cut -c1-300
