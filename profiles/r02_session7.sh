#!/bin/bash
# Round 2, GPU session 7: variant 10 (pipelined z-march) parity and timings; full suite with variant 9 as the default.
OUT=gpurun_out/r02h
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "zpipe" > $OUT/pytest_zpipe.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_zpipe.log; tail -6 $OUT/pytest_zpipe.log
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$name.json"))
    print("$name", round(d["ms_per_step"],3), "ms", round(d["value"],1), "MLUPS", d["roofline"]["kernel_ms"], d["clocks"])
except Exception as e:
    print("$name failed", e, open("$OUT/bench_$name.err").read()[-600:])
PY
}
run v9 MBL_VARIANT=9
run v10_w8 MBL_VARIANT=10
run v10_w4 MBL_VARIANT=10 MBL_ROWS=4
run v10_w8_zm16 MBL_VARIANT=10 MBL_ZMARCH=16
run v10_w8_zm4 MBL_VARIANT=10 MBL_ZMARCH=4
MBL_VARIANT=10 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'k_collide|k_qcorr' -s 6 -c 4 --csv --log-file $OUT/traffic_512_zpipe.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/traffic_512_zpipe.log 2>&1
tail -5 $OUT/traffic_512_zpipe.csv | cut -c1-60,170-400
timeout 2400 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
