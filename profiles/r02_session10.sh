#!/bin/bash
# Round 2, GPU session 10 (2 GPUs): the z-march kernels inside the overlapped split step -- parity on one device, the
# 2-rank check on real NCCL transport, weak / strong timings against the blocking exchange.
N=2
OUT=gpurun_out/r02n
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "overlapped or slab or lean_halo" > $OUT/pytest_slabs.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_slabs.log; tail -5 $OUT/pytest_slabs.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 profiles/n2check.py > $OUT/n2check.txt 2>&1
tail -9 $OUT/n2check.txt
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
run weak_lean_overlap NCCL_DEBUG=WARN -- --scaling weak
run weak_lean_overlap_v5 MBL_VARIANT=5 -- --scaling weak
run weak_blocking MBL_OVERLAP=0 -- --scaling weak
run strong_lean_overlap NCCL_DEBUG=WARN -- --scaling strong
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > $OUT/n1.json 2> $OUT/n1.err; python -c "
import json; d=json.load(open('$OUT/n1.json')); print('n1', d['ms_per_step'], d['value'])"
