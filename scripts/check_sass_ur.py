"""Static sanity check of the built library: flags kernels in which an asynchronous-copy instruction (LDGSTS, UBLKCP)
READS a uniform register that no instruction of the kernel WRITES.  Seen once with ptxas 12.9: the 2-quads-per-lane
cp.async instantiations of k_block_stg came out as `LDGSTS [R0+UR0], desc[UR1]` with UR0/UR1 never defined and
faulted with 'illegal instruction' at run time (compute-sanitizer).  Usage: python scripts/check_sass_ur.py [lib.so]"""
import re
import subprocess
import sys


def check(lib):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    bad = []
    fns = txt.split("Function : ")[1:]
    for fn in fns:
        name = fn.split("\n", 1)[0].strip()
        written, need = set(), set()
        for line in fn.splitlines():
            m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);", line)
            if not m:
                continue
            ins = re.sub(r"^@!?U?P\d+\s+", "", m.group(1).strip())
            parts = ins.split(None, 1)
            if len(parts) < 2:
                continue
            op, args = parts
            if op.startswith(("LDGSTS", "UBLKCP")):
                for r in re.findall(r"\bUR(\d+)\b", args):
                    need.add(int(r))
                continue
            first = args.split(",")[0]
            d = re.findall(r"\bUR(\d+)\b", first)
            if d and (op.startswith("U") or op.startswith(("LDCU", "S2UR", "R2UR", "VOTEU", "REDUX"))):
                r = int(d[0])
                n = 4 if ".128" in op else (2 if (".64" in op or "WIDE" in op) else 1)
                for i in range(n):
                    written.add(r + i)
        undefined = sorted(r for r in need if r not in written)
        if undefined:
            bad.append((name, undefined))
    return len(fns), bad


if __name__ == "__main__":
    n, bad = check(sys.argv[1] if len(sys.argv) > 1 else "relationprediction_b200/lib/librgcn_b200.so")
    for name, regs in bad:
        print("async copy reads undefined uniform registers %s in %s" % (regs, name[:160]))
    print("checked %d kernels, %d suspicious" % (n, len(bad)))
    sys.exit(1 if bad else 0)
