#!/usr/bin/env python
"""profiles/r2_traffic.json from an `ncu --page raw --csv` dump holding dram__bytes_read.sum / dram__bytes_write.sum of
the aggregation kernels:  python scripts/make_traffic_json.py <raw.csv> "<workload>|<scale>" [more csv/key pairs]
bench.py reads the entry of its workload and flags it stale when the library stamp differs."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (lib_stamp)


def stage_of(kernel):
    if "k_block_team" in kernel or "k_block_stg" in kernel or "k_block_rel" in kernel:
        tail = kernel[kernel.index("k_block_"):]
        args = tail.split("<", 1)[1].split(">")[0].replace("(int)", "").replace("(bool)", "").split(",")
        fused = args[2].strip() in ("1", "true")
        return "block_agg_dH" if fused else "block_agg_fwd"
    return None


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main():
    out_p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    tj = json.load(open(out_p)) if os.path.exists(out_p) else {}
    for src, key in zip(sys.argv[1::2], sys.argv[2::2]):
        rows = list(csv.reader(open(src)))
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        ent = {"lib_stamp": bench.lib_stamp(), "kernel_source_stamp": bench.kernel_source_stamp(),
               "source": os.path.basename(src)}
        for r in rows[2:]:
            st = stage_of(r[ix["Kernel Name"]])
            if st is None or st in ent:
                continue
            rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
            wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
            ent[st] = int(rd + wr)
            ent[st + "_read_write"] = [int(rd), int(wr)]
        tj[key] = ent
    json.dump(tj, open(out_p, "w"), indent=1, sort_keys=True)
    print(json.dumps(tj, indent=1))


if __name__ == "__main__":
    main()
