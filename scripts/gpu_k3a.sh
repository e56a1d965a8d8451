#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== box after own cell"; python scripts/stage_times.py 512 3 2>&1 | tail -7
echo "== shells only"; GICP_BOX_AFTER_OWN=0 python scripts/stage_times.py 512 3 2>&1 | tail -7
