#!/bin/bash
# Round 2, GPU session 1: parity of the march step (variant 8), the value-level pin at 256^3 / 512^3, the plain-C
# caller on a device, then step timings at 512^3: tile carry step (5) against the march step in several tunings,
# and the DRAM traffic of the march kernel.
OUT=gpurun_out/r02a
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,power.limit --format=csv > $OUT/gpu.csv 2>&1
nproc > $OUT/host.txt; free -g >> $OUT/host.txt
timeout 1200 python -m pytest tests -m gpu -x -q -k "march or tiled or c_caller or test_cabi" > $OUT/pytest_march.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_march.log; tail -5 $OUT/pytest_march.log
run() { # name, env...
  name=$1; shift
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
run v8_zm64 MBL_VARIANT=8
run v8_zm64_nopipe MBL_VARIANT=8 MBL_PIPE=0
run v8_zm32 MBL_VARIANT=8 MBL_ZM=32
run v8_zm128 MBL_VARIANT=8 MBL_ZM=128
run v8_zm512 MBL_VARIANT=8 MBL_ZM=512
run v0 MBL_VARIANT=0
MBL_VARIANT=8 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:'k_march' -s 3 -c 2 --csv --log-file $OUT/traffic_512_march.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/traffic_512_march.log 2>&1
tail -4 $OUT/traffic_512_march.csv | cut -c1-80,180-500
MBL_VARIANT=8 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_march' -s 3 -c 1 -o $OUT/ncu_full_march_256 \
    python bench.py --size 256 --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full_march_256.log 2>&1
ls -la $OUT
