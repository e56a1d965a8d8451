#!/bin/bash
for p in 16 32 64; do echo "== acc ppt $p"; GICP_ACC_PPT=$p python scripts/stage_times.py 1024 3 2>&1 | tail -7; done
