#!/bin/bash
# One GPU session: parity tests, bench (both arms), ncu launch list, ncu full capture of the two hot kernels.
# Usage (from the repo root, under gpurun): bash profiles/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.csv 2>&1
nproc > $OUT/host.txt; free -g >> $OUT/host.txt
if [ -z "$SKIP_TESTS" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
  tail -3 $OUT/pytest.log
fi
timeout 600 python bench.py --impl reference > $OUT/bench_reference.json 2> $OUT/bench_reference.err; cat $OUT/bench_reference.json
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; cat $OUT/bench.json; tail -3 $OUT/bench.err
if [ -z "$SKIP_NCU" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_512.csv \
      python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/launches_512.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_collide|k_qcorr' -s 6 -c 2 -f -o $OUT/prof_256 \
      python bench.py --size 256 --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/prof_256.log 2>&1
fi
ls -la $OUT
