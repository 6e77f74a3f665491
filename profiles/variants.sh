#!/bin/bash
# timings of the fused-through-L2 variants (1: TMA persistent, 3: plain persistent) against the default two-kernel step,
# and their DRAM / L2 traffic (ncu, 256^3)
OUT=gpurun_out/${1:-variants}
mkdir -p $OUT
for v in 0 1 3; do
  timeout 300 python bench.py --variant $v --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_v$v.json 2> $OUT/bench_v$v.err
  cat $OUT/bench_v$v.json
done
for v in 1 3; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
     --clock-control none -k regex:'k_fused' -s 3 -c 1 --csv --log-file $OUT/ncu_v$v.csv \
     python bench.py --variant $v --size 256 --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_v$v.log 2>&1
done
