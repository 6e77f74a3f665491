#!/bin/bash
# Round 2, multi-GPU session (N = $1 GPUs of one box): strong scaling of the 512^3 Taylor-Green box cut into z-slabs
# (lean z-halo, exchange hidden behind the interior planes) against the blocking full-halo exchange; at N = 2 also the
# weak-scaling line and the wall / EB workload with the overlapped exchange on and off.
N=$1
OUT=gpurun_out/r02m$N
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
run() { # name, extra env..., -- bench args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
      bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-cpu "$@" > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/$name.json") if l.startswith("{")][-1])
    print("$name", round(d["ms_per_step"],3), "ms", round(d["value"],1), "MLUPS", d["scaling"], d["config"]["decomposition"], d["roofline"]["kernel_ms"], d["clocks"]["reasons"])
except Exception as e:
    print("$name failed", e, open("$OUT/$name.err").read()[-800:])
PY
}
run strong_lean_overlap NCCL_DEBUG=WARN -- --scaling strong
run strong_full_blocking MBL_HALO_LEAN=0 MBL_OVERLAP=0 -- --scaling strong
run strong_full_overlap MBL_HALO_LEAN=0 -- --scaling strong
if [ "$N" = "2" ]; then
  run weak_lean_overlap NCCL_DEBUG=WARN -- --scaling weak
  run weak_full_overlap MBL_HALO_LEAN=0 -- --scaling weak
  run channel_overlap NCCL_DEBUG=WARN -- --workload channel
  run channel_blocking MBL_OVERLAP=0 -- --workload channel
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 profiles/n2check.py > $OUT/n2check.txt 2>&1
  tail -3 $OUT/n2check.txt
fi
if [ "$N" = "8" ]; then
  run weak_lean_overlap NCCL_DEBUG=WARN -- --scaling weak
fi
ls $OUT
