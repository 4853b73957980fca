#!/usr/bin/env python
"""Write profiles/r2_sass_evidence.txt (+ per-kernel listings): which Blackwell-native instructions each kernel of the
built library contains (cuobjdump -sass).  Run here after build(); no GPU needed."""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "relationprediction_b200", "lib", "librgcn_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "REDG", "FFMA", "SHFL", "BAR"]
LISTINGS = {   # demangled-name substring -> file
    "k_block_team<8, 1, false, false": "r2_sass_k_block_team_s8_fwd.txt",
    "k_block_team<8, 1, true, false": "r2_sass_k_block_team_s8_fused_bwd.txt",
    "k_gemm_tf32x3<0>": "r2_sass_k_gemm_tf32x3_nt.txt",
    "k_gemm_tn_tf32x3": "r2_sass_k_gemm_tf32x3_tn.txt",
}

def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    rows = []
    for fn in txt.split("Function : ")[1:]:
        name = fn.split("\n", 1)[0].strip()
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = dem.replace("(anonymous namespace)::", "").replace("void ", "")
        short = re.sub(r"\(.*", "", dem)
        c, body = collections.Counter(), []
        for line in fn.splitlines():
            m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?)\s*/\* 0x", line)
            if not m:
                continue
            ins = m.group(2)
            body.append(m.group(1) + "  " + ins)
            op = re.sub(r"^@!?U?P\d+\s+", "", ins)
            for k in KEYS:
                if op.startswith(k):
                    c[k] += 1
        rows.append((short, len(body), c))
        for sub, fname in LISTINGS.items():
            if sub in dem:
                with open(os.path.join(ROOT, "profiles", fname), "w") as f:
                    f.write("# %s\n" % dem)
                    f.write("\n".join(body) + "\n")
    with open(os.path.join(ROOT, "profiles", "r2_sass_evidence.txt"), "w") as f:
        f.write("# cuobjdump -sass relationprediction_b200/lib/librgcn_b200.so: instruction mix per kernel (scripts/sass_evidence.py)\n")
        f.write("# UTCHMMA = tcgen05.mma (kind::tf32), LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (TMA bulk copy),\n")
        f.write("# LDGSTS = cp.async, SYNCS = mbarrier ops, REDG = red.global.add(.v4).f32\n\n")
        for short, n, c in sorted(rows):
            f.write("%-78s %5d instr  %s\n" % (short[:78], n, "  ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k])))

if __name__ == "__main__":
    sys.exit(main())
