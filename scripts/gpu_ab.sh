#!/bin/bash
# A/B on the GPU box: tests, then per-stage times of 512 pairs under the table-size switch
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== new default"; python scripts/stage_times.py 512 3 2>&1 | tail -7
echo "== old table (2^30)"; GICP_CELL_TABLE_LOG2=30 python scripts/stage_times.py 512 3 2>&1 | tail -7
echo "== table 2^24 (2^15 cells per cloud)"; GICP_CELL_TABLE_LOG2=24 python scripts/stage_times.py 512 3 2>&1 | tail -7
