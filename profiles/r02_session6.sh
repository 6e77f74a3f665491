#!/bin/bash
# Round 2, GPU session 6: the z-march carry step (variant 9): parity, then timings at 512^3 against the default.
OUT=gpurun_out/r02g
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q -k "zmarch" > $OUT/pytest_zmarch.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_zmarch.log; tail -6 $OUT/pytest_zmarch.log
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
run v5 MBL_VARIANT=5
run v9_zm8 MBL_VARIANT=9
run v9_zm4 MBL_VARIANT=9 MBL_ZMARCH=4
run v9_zm16 MBL_VARIANT=9 MBL_ZMARCH=16
run v9_zm32 MBL_VARIANT=9 MBL_ZMARCH=32
run v7 MBL_VARIANT=7
MBL_VARIANT=9 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'k_collide|k_qcorr' -s 6 -c 4 --csv --log-file $OUT/traffic_512_zmarch.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/traffic_512_zmarch.log 2>&1
tail -12 $OUT/traffic_512_zmarch.csv | cut -c1-60,170-400
