"""Turns the raw ncu outputs a gpurun call left in gpurun_out/ into the small, committed summaries
under profiles/ (launch list shares + per-kernel metric table)."""
import collections
import csv
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
# ---- launch list ----
rows = list(csv.reader(open("gpurun_out/launches_%s.csv" % tag)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0][:90]
    v = float(r[vi].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r[ui], v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
with open("profiles/%s_launches.txt" % tag, "w") as fh:
    fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 2 --warmup 3 --no-cpu-baseline\n")
    fh.write("# per-kernel totals over the whole process (cold-cache, serialised: compare SHARES)\n")
    fh.write("%12s %7s %6s  %s\n" % ("total_us", "share", "count", "kernel"))
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        fh.write("%12.1f %6.1f%% %6d  %s\n" % (t, 100 * t / tot, n, k))
# ---- full capture ----
rows = list(csv.reader(open("gpurun_out/prof_%s_raw.csv" % tag)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
with open("profiles/%s_ncu_kernels.md" % tag, "w") as fh:
    fh.write("# ncu --set full --clock-control none: one launch of each hot kernel, default bench (FB15k-237 shape)\n\n")
    seen = set()
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        if name in seen:
            continue
        seen.add(name)
        fh.write("## %s\n\n| metric | value | unit |\n|---|---|---|\n" % name)
        for w in want:
            if w in idx:
                fh.write("| %s | %s | %s |\n" % (w, r[idx[w]], units[idx[w]]))
        st = [(h, float(r[i].replace(",", ""))) for i, h in enumerate(hdr)
              if "issue_stalled" in h and h.endswith("_per_warp_active.pct") and r[i] not in ("", "n/a")]
        for h, v in sorted(st, key=lambda x: -x[1])[:5]:
            fh.write("| stall: %s | %.2f | %% of warp-active |\n" % (h.split("issue_stalled_")[1].replace("_per_warp_active.pct", ""), v))
        fh.write("\n")
print("profiles written")
