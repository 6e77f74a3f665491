#!/bin/bash
# Round 2, GPU session 2: the multi-box / multi-level (AMR-exact) mode against the reference goldens and the
# multi-level oracle, then the whole GPU suite (the single-box kernels were refactored around it).
OUT=gpurun_out/r02b
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_amr.py -x -q -s > $OUT/pytest_amr.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_amr.log; tail -40 $OUT/pytest_amr.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_amr.py > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -8 $OUT/pytest_gpu.log
