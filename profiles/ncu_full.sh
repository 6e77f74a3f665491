#!/bin/bash
# one ncu --set full capture of a kernel at 256^3:  bash profiles/ncu_full.sh <tag> <kernel-regex> <variant> [ENV=VAL ...]
TAG=$1; KREGEX=$2; V=$3; shift 3
OUT=gpurun_out/$TAG
mkdir -p $OUT
env "$@" timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s 3 -c 1 -f -o $OUT/prof \
    python bench.py --variant $V --size 256 --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/prof.log 2>&1
tail -2 $OUT/prof.log
ls -la $OUT
