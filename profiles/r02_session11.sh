#!/bin/bash
# Round 2, GPU session 11: compute-sanitizer over the new default step (z-march kernels, in the split step too) and the
# new AMR entry points; then the full GPU suite, smoke() and both arms of bench.py on the final tree.
OUT=gpurun_out/r02o
mkdir -p $OUT
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 $SAN --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests -m gpu -x -q \
    -k "(golden and (chcyl or touch or tg12) and (zmarch or default or unfused)) or odd_box or amr_cuda_vs_reference_golden or (overlapped and (chcyl or pressure)) or graph_replay" \
    > $OUT/memcheck.log 2>&1
echo "memcheck rc=$?" >> $OUT/memcheck.log; tail -4 $OUT/memcheck.log
timeout 1500 $SAN --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests -m gpu -x -q \
    -k "(golden and (chcyl or tg12) and zmarch) or (odd_box and zmarch) or (overlapped and tg12 and zmarch)" \
    > $OUT/racecheck.log 2>&1
echo "racecheck rc=$?" >> $OUT/racecheck.log; tail -4 $OUT/racecheck.log
timeout 2400 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log; tail -5 $OUT/smoke.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_reference_n1.json 2> $OUT/bench_reference_n1.err; cut -c1-400 $OUT/bench_reference_n1.json
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; cut -c1-1500 $OUT/bench_n1.json
