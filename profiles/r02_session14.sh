#!/bin/bash
# Round 2, GPU session 14 (8 GPUs): weak and strong scaling of the default (z-march) step, the distributed 2-level channel.
N=8
OUT=gpurun_out/r02t
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo_n8.txt 2>&1
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
run n8_weak NCCL_DEBUG=WARN -- --scaling weak
run n8_strong NCCL_DEBUG=WARN -- --scaling strong
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu > $OUT/n1.json 2> $OUT/n1.err; python -c "
import json; d=json.load(open('$OUT/n1.json')); print('n1', d['ms_per_step'], d['value'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29516 profiles/amr_bench.py --nx 1024 --ny 256 --nz 64 --steps 5 --warmup 2 > $OUT/amr_1024_n8.json 2> $OUT/err3.txt; grep "^{" $OUT/amr_1024_n8.json; tail -2 $OUT/err3.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 profiles/amr_n2check.py > $OUT/amr_n8check.txt 2>&1; tail -6 $OUT/amr_n8check.txt
