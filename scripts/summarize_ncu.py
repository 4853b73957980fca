"""Turns an `ncu --page raw --csv` dump (gpurun_out/<name>_raw.csv) into a committed markdown summary under profiles/:
   python scripts/summarize_ncu.py gpurun_out/r2_prof_syn_raw.csv profiles/r2_ncu_synthetic.md "title line" """
import csv
import sys

src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "lts__t_sectors_srcunit_tex_op_red.sum"]
with open(dst, "w") as fh:
    fh.write("# %s\n\n" % title)
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
              if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")]
        for h, v in sorted(st, key=lambda x: -x[1])[:6]:
            fh.write("| stalled warps per issue cycle: %s | %.2f | warps |\n"
                     % (h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), v))
        fh.write("\n")
print("wrote", dst)
