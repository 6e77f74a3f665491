#!/bin/bash
# carry step (variant 4): parity tests, then timing sweep over its tunings against the two-kernel step (variant 0),
# then DRAM traffic of its two kernels (ncu, 256^3)
OUT=gpurun_out/${1:-carry}
mkdir -p $OUT
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "carry or tile" > $OUT/pytest_carry.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_carry.log
  tail -5 $OUT/pytest_carry.log
fi
run() { # name, variant, env...
  name=$1; v=$2; shift 2
  env "$@" timeout 300 python bench.py --variant $v --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python - "$name" "$OUT/bench_$name.json" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); print(sys.argv[1], "ms/step %.2f MLUPS %.0f" % (d["ms_per_step"], d["value"]), d["roofline"]["kernel_ms"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run v0 0 X=1
# each config: comma-separated environment assignments (MBL_KY rows per march, MBL_OWN cells per warp, MBL_MINB CTAs/SM,
# MBL_SYNC CTA barrier per row, MBL_PREFETCH rows of L2 prefetch distance)
for cfg in ${CONFIGS:-"MBL_MINB=3" "MBL_MINB=2"}; do
  v=${VARIANT:-4}
  run v${v}_$(echo $cfg | sed 's/MBL_//g' | tr ',=' '__') $v $(echo $cfg | tr ',' ' ')
done
if [ -z "$SKIP_NCU" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum \
     --clock-control none -k regex:'k_collide_carry|k_collide_tile|k_qcorr_combine' -s 4 -c 2 --csv --log-file $OUT/ncu_carry_256.csv \
     python bench.py --variant ${VARIANT:-4} --size 256 --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_carry_256.log 2>&1
  tail -4 $OUT/ncu_carry_256.csv | cut -c1-400
fi
