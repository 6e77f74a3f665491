#!/bin/bash
# Round 2, GPU session 8: where do the 5 ms between k_collide_lean (20.6 ms) and k_collide_tile_march (25.8 ms) go?
# Timing-only ablations of the z-march kernel (MBL_EXPERIMENTS build; MBL_ABLATE bits: 1 no x shuffles, 2 no y exchange
# / barriers, 4 no carried stores -> the sums are dead code; MBL_OWN=32: no halo lanes).  Results of ablated runs are wrong.
OUT=gpurun_out/r02j
mkdir -p $OUT
export MBL_EXPERIMENTS=1
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$name.json"))
    print("$name", round(d["ms_per_step"],3), "ms", d["roofline"]["kernel_ms"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name failed", e, open("$OUT/bench_$name.err").read()[-600:])
PY
}
run v0 MBL_VARIANT=0
run v9 MBL_VARIANT=9
run v9_noshfl MBL_VARIANT=9 MBL_ABLATE=1
run v9_nobar MBL_VARIANT=9 MBL_ABLATE=2
run v9_nostores MBL_VARIANT=9 MBL_ABLATE=4
run v9_nocarry MBL_VARIANT=9 MBL_ABLATE=6
run v9_nocarry_own32 MBL_VARIANT=9 MBL_ABLATE=6 MBL_OWN=32
run v9_own32 MBL_VARIANT=9 MBL_OWN=32
run v9_again MBL_VARIANT=9
