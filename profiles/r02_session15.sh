#!/bin/bash
# Round 2, GPU session 15: final validation of the tree -- full GPU suite, smoke(), both arms of bench.py, the wall / EB
# workload with the default step.
OUT=gpurun_out/r02z
mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log; tail -4 $OUT/smoke.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_reference_n1.json 2> $OUT/bench_reference_n1.err; cut -c1-300 $OUT/bench_reference_n1.json
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; cut -c1-1300 $OUT/bench_n1.json
timeout 600 python bench.py --workload channel --steps 20 --warmup 3 --no-e2e --no-cpu > $OUT/bench_channel_n1.json 2> $OUT/bench_channel_n1.err; python -c "
import json; d=json.load(open('$OUT/bench_channel_n1.json')); print('channel', d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['config']['variant'][:40])"
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > $OUT/gpu.csv
