#!/bin/bash
# Round 2, GPU session 16: order in which k_collide_tile_march issues a plane's loads (experiments build, MBL_LOADORDER):
# 0 g (cp.async), f, mask / QCorr (shipped); 1 f, mask / QCorr, g; 2 mask / QCorr, f, g.  Same results, same box.
OUT=gpurun_out/r02o2
mkdir -p $OUT
export MBL_EXPERIMENTS=1
for o in 0 1 2 0; do
  MBL_LOADORDER=$o timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_ord$o.json 2> $OUT/err.txt
  python -c "
import json; d=json.load(open('$OUT/bench_ord$o.json')); print('order $o', round(d['ms_per_step'],3), d['roofline']['kernel_ms'], d['clocks']['sm_mhz'])"
done
