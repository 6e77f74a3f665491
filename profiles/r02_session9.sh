#!/bin/bash
# Round 2, GPU session 9: the AMR-exact mode at size (channel + EB cylinder, 2 levels), its kernel shares, the reference
# CPU run of the same deck.
OUT=gpurun_out/r02k
mkdir -p $OUT
timeout 600 python profiles/amr_bench.py --steps 10 --warmup 3 > $OUT/amr_512.json 2> $OUT/amr_512.err; tail -2 $OUT/amr_512.err; cat $OUT/amr_512.json
timeout 600 python profiles/amr_bench.py --nx 1024 --ny 256 --nz 64 --mgs 64 --steps 5 --warmup 2 > $OUT/amr_1024.json 2> $OUT/amr_1024.err; tail -2 $OUT/amr_1024.err; cat $OUT/amr_1024.json
timeout 600 python profiles/amr_bench.py --mgs 32 --steps 10 --warmup 3 > $OUT/amr_512_mgs32.json 2> $OUT/amr_512_mgs32.err; tail -2 $OUT/amr_512_mgs32.err; cat $OUT/amr_512_mgs32.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_amr.csv \
    python profiles/amr_bench.py --steps 1 --warmup 1 > $OUT/launches_amr.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/r02k/launches_amr.csv") if l.startswith('"')))
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    name = r[ki].split("(")[0]
    tot[name] += float(r[vi].replace(",", "")) / 1e6; cnt[name] += 1
allms = sum(tot.values())
with open("gpurun_out/r02k/launches_amr_summary.txt", "w") as fh:
    fh.write(f"ncu launch list of profiles/amr_bench.py --steps 1 --warmup 1 (setup + init + 2 coarse steps), ms under ncu\n")
    for k, v in tot.most_common(25):
        fh.write(f"{v:10.3f} ms {100 * v / allms:5.1f} %  x{cnt[k]:5d}  {k}\n")
print(open("gpurun_out/r02k/launches_amr_summary.txt").read())
PY
timeout 900 python profiles/amr_bench.py --reference --ref-steps 3 > $OUT/amr_512_reference.json 2> $OUT/amr_512_reference.err; cat $OUT/amr_512_reference.json
