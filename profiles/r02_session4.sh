#!/bin/bash
# Round 2, GPU session 4: the whole GPU suite (device regrid, device geometry, overlapped slabs with walls, new goldens),
# smoke, and the launch list + DRAM traffic of the wall / EB workload.
OUT=gpurun_out/r02d
mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -12 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 60 -c 48 --csv \
    --log-file $OUT/launches_channel.csv python bench.py --workload channel --steps 4 --warmup 3 --no-cpu > $OUT/launches_channel.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02d/launches_channel.csv")) if len(r) > 10 and r[0].isdigit()]
t = collections.defaultdict(float); n = collections.Counter(); by = collections.defaultdict(float)
for r in rows:
    name = r[4].split("(")[0][-60:]
    if r[-3] == "gpu__time_duration.sum": t[name] += float(r[-1]); n[name] += 1
    elif r[-3].startswith("dram__bytes"): by[name] += float(r[-1])
tot = sum(t.values())
with open("gpurun_out/r02d/launches_channel_summary.txt", "w") as fh:
    for k, v in sorted(t.items(), key=lambda kv: -kv[1])[:14]:
        line = f"{k:62s} {n[k]:4d} launches {v/1e6:9.3f} ms {100*v/tot:5.1f} %  {by[k]/1e9:8.2f} GB"
        print(line); fh.write(line + "\n")
PY
ls $OUT
