#!/bin/bash
OUT=gpurun_out/r02e
mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -15 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log
