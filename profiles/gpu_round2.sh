#!/bin/bash
# Round-1 (session 7) GPU evidence: full parity suite, bench (both arms), ncu launch list at 512^3, DRAM traffic of
# the two hot kernels at 512^3 (one pass, no replay), full ncu capture at 256^3.
TAG=${1:-r01b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.csv 2>&1
nproc > $OUT/host.txt; free -g >> $OUT/host.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log
timeout 600 python bench.py --impl reference > $OUT/bench_reference.json 2> $OUT/bench_reference.err; cat $OUT/bench_reference.json | cut -c1-300
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; cat $OUT/bench.json; tail -3 $OUT/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_512.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/launches_512.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'k_collide|k_qcorr' -s 6 -c 2 --csv --log-file $OUT/traffic_512.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/traffic_512.log 2>&1
grep -c . $OUT/traffic_512.csv; tail -6 $OUT/traffic_512.csv | cut -c1-60,200-400
ls -la $OUT
