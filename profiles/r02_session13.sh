#!/bin/bash
# Round 2, GPU session 13 (2 GPUs): distributed multi-box levels on real NCCL transport -- bit-identity against one rank,
# and the 2-level channel at size on 1 and 2 GPUs.
OUT=gpurun_out/r02s
mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 profiles/amr_n2check.py > $OUT/amr_n2check.txt 2>&1
echo "rc=$?" >> $OUT/amr_n2check.txt; tail -8 $OUT/amr_n2check.txt
timeout 300 python profiles/amr_bench.py --steps 10 --warmup 3 > $OUT/amr_512_n1.json 2> $OUT/err1.txt; cat $OUT/amr_512_n1.json; tail -2 $OUT/err1.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 profiles/amr_bench.py --steps 10 --warmup 3 > $OUT/amr_512_n2.json 2> $OUT/err2.txt; grep "^{" $OUT/amr_512_n2.json; tail -3 $OUT/err2.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 profiles/amr_bench.py --nx 1024 --ny 256 --nz 64 --steps 5 --warmup 2 > $OUT/amr_1024_n2.json 2> $OUT/err3.txt; grep "^{" $OUT/amr_1024_n2.json; tail -3 $OUT/err3.txt
timeout 300 python profiles/amr_bench.py --nx 1024 --ny 256 --nz 64 --steps 5 --warmup 2 > $OUT/amr_1024_n1.json 2> $OUT/err4.txt; cat $OUT/amr_1024_n1.json
