#!/usr/bin/env bash
# where the second resize mode (tile path) spends its time: wall clock, then a per-kernel launch list under ncu
set -uo pipefail
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python bench.py --config antialias --images 4 --steps 2 --warmup 1 2>gpurun_out/aa.err | tail -1 | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/aa_launches.csv python bench.py --config antialias --images 2 --steps 1 --warmup 1 > gpurun_out/aa_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/aa_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    v = v / 1e3 if u in ("ns", "nsecond") else v * (1e3 if u.startswith("ms") else 1)   # -> us
    tot[r[ki][:70]] += v; cnt[r[ki][:70]] += 1
s = sum(tot.values())
print(f"total kernel time {s/1e3:.1f} ms over {sum(cnt.values())} launches")
for k, v in tot.most_common(14): print(f"{v/1e3:9.2f} ms {100*v/s:5.1f}% {cnt[k]:6d}  {k}")
PY
rm -f gpurun_out/aa_launches.csv
