"""Aggregate an ncu --metrics gpu__time_duration.sum --csv launch list by kernel name."""
import collections, csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i
        break
ik, iv = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) <= iv:
        continue
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    name = r[ik].split("(")[0].replace("unmore::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':58s} {'launches':>8s} {'total ms':>10s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:58]:58s} {v[0]:8d} {v[1]/1e6:10.3f} {100*v[1]/tot:6.2f}%")
print(f"{'TOTAL':58s} {sum(v[0] for v in agg.values()):8d} {tot/1e6:10.3f}")
if len(sys.argv) > 2:
    d = json.load(open(sys.argv[2]))
    print("\nlive CUDA-event shares from the plain run of the same command (bench.py `kernels`):")
    for k, v in d["kernels"].items():
        print(f"  {k:34s} {v['ms_per_step']:10.3f} ms  {100*v['share']:6.2f}%")
    print(f"  step: {d['ms_per_step']:.3f} ms, {d['value']:.1f} images/s, gpu_launches {d['gpu_launches']}")
