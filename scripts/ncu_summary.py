"""Summarise an .ncu-rep: key raw metrics + executed-instruction opcode histogram (needs ncu on PATH)."""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
norm = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0   # divide instruction counts by this (e.g. rounds*128)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
pat = re.compile(r"^(gpu__time_duration.sum|launch__registers_per_thread|launch__grid_size|launch__block_size|"
                 r"sm__warps_active.avg.pct_of_peak_sustained_active|sm__throughput.avg.pct_of_peak_sustained_elapsed|"
                 r"gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|dram__bytes_read.sum|dram__bytes_write.sum|"
                 r"smsp__inst_executed.sum|smsp__issue_active.avg.pct_of_peak_sustained_active|"
                 r"sm__inst_executed_pipe_(alu|fma|xu|lsu|fp64).avg.pct_of_peak_sustained_active|"
                 r"l1tex__t_sector_hit_rate.pct|lts__t_sector_hit_rate.pct|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum|"
                 r"smsp__average_warps_issue_stalled_(long_scoreboard|short_scoreboard|wait|not_selected|math_pipe_throttle|mio_throttle|branch_resolving|barrier|dispatch_stall|lg_throttle)_per_issue_active.ratio)$")
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")])
    for i, h in enumerate(hdr):
        if pat.match(h):
            print(f"  {h} [{units[i]}] = {r[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
isrc, iex = hdr.index("Source"), hdr.index("Instructions Executed")
hist = collections.Counter()
tot = 0
for r in rows[2:]:
    try:
        n = int(r[iex])
    except Exception:
        continue
    toks = r[isrc].strip().split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    hist[op.split(".")[0]] += n
    tot += n
print(f"instructions executed: {tot}  (normalised /{norm:g} = {tot/norm:.1f})")
for k, v in hist.most_common(24):
    print(f"  {k:10s} {100*v/tot:6.2f}%  {v/norm:9.2f}")
