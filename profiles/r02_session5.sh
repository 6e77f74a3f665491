#!/bin/bash
# Round 2, GPU session 5: compute-sanitizer (memcheck, racecheck on shared memory) over the small-deck parity tests of the
# shipped step variants and of the AMR mode; one full ncu capture of the default collide kernel on the wall / EB workload.
OUT=gpurun_out/r02f
mkdir -p $OUT
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 $SAN --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests -m gpu -x -q \
    -k "(golden and (chcyl or touch or slipyz or tg12) and (tile or twopass or unfused or pair)) or amr_cuda_vs_reference_golden or regrid or lean_halo or (overlapped and chcyl)" \
    > $OUT/memcheck.log 2>&1
echo "memcheck rc=$?" >> $OUT/memcheck.log; tail -6 $OUT/memcheck.log
timeout 1500 $SAN --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests -m gpu -x -q \
    -k "(golden and (chcyl or tg12) and (tile or pair or twopass)) or (amr_cuda_vs_reference_golden and amr2_chcyl)" \
    > $OUT/racecheck.log 2>&1
echo "racecheck rc=$?" >> $OUT/racecheck.log; tail -6 $OUT/racecheck.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_collide_tile' -s 4 -c 1 -o $OUT/ncu_full_tile6_channel \
    python bench.py --workload channel --size 256 --steps 2 --warmup 3 --no-cpu > $OUT/ncu_full_tile6_channel.log 2>&1
ls -la $OUT
