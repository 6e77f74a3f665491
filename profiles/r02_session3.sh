#!/bin/bash
# Round 2, GPU session 3: remaining new tests (AMR random state / bind / regrid, lean halo), drift on the hard decks,
# the bench line with the reworked reference arm, and the wall / EB workload with its launch list and traffic.
OUT=gpurun_out/r02c
mkdir -p $OUT
nproc > $OUT/host.txt; free -g >> $OUT/host.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "amr or lean or slab or overlapped or tiled" > $OUT/pytest_new.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_new.log; tail -6 $OUT/pytest_new.log
timeout 600 python profiles/drift.py > $OUT/drift.jsonl 2> $OUT/drift.err; cut -c1-260 $OUT/drift.jsonl; tail -3 $OUT/drift.err
timeout 600 python bench.py --workload channel --steps 20 --warmup 3 --no-cpu > $OUT/bench_channel.json 2> $OUT/bench_channel.err
cut -c1-900 $OUT/bench_channel.json; tail -3 $OUT/bench_channel.err
timeout 600 python bench.py --workload channel --steps 20 --warmup 3 --no-cpu --variant 0 > $OUT/bench_channel_v0.json 2> $OUT/bench_channel_v0.err
cut -c1-400 $OUT/bench_channel_v0.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 200 -c 120 --csv \
    --log-file $OUT/launches_channel.csv python bench.py --workload channel --steps 3 --warmup 3 --no-cpu > $OUT/launches_channel.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02c/launches_channel.csv")) if len(r) > 10 and r[0].isdigit()]
t = collections.defaultdict(float); n = collections.Counter(); by = collections.defaultdict(float)
for r in rows:
    name = r[4].split("(")[0][-60:]
    if r[-3] == "gpu__time_duration.sum": t[name] += float(r[-1]); n[name] += 1
    elif r[-3].startswith("dram__bytes"): by[name] += float(r[-1])
tot = sum(t.values())
for k, v in sorted(t.items(), key=lambda kv: -kv[1])[:12]:
    print(f"{k:62s} {n[k]:4d} launches {v/1e6:9.3f} ms {100*v/tot:5.1f} %  {by[k]/1e9:8.2f} GB")
PY
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; cut -c1-1500 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference_n1.json 2> $OUT/bench_reference_n1.err; cut -c1-1200 $OUT/bench_reference_n1.json
ls -la $OUT
