"""Per-kernel SASS hot spots from an .ncu-rep: opcode histogram of executed warp instructions and the instructions
with the most stall samples.   usage: python profiles/ncu_sass.py REP [kernel-substring] [top]"""
import csv
import io
import subprocess
import sys
from collections import Counter


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": [], "hdr": None}
            kernels.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] is not None and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    seen = set()
    for k in kernels:
        if want not in k["name"] or k["name"] in seen:
            continue
        seen.add(k["name"])
        h = k["hdr"]
        si, ni, ei = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        tot_s = sum(int(r[ni]) for r in k["rows"]) or 1
        tot_e = sum(int(r[ei]) for r in k["rows"]) or 1
        print(f"== {k['name']}: {len(k['rows'])} SASS instrs, {tot_e} warp-instr executed, {tot_s} samples")
        ops = Counter()
        for r in k["rows"]:
            op = r[si].split()
            op = [t for t in op if not t.startswith("@")][0].split(".")[0] if op else "?"
            ops[op] += int(r[ei])
        print("   executed by opcode:", ", ".join(f"{o} {100.0 * n / tot_e:.1f}%" for o, n in ops.most_common(22)))
        order = sorted(range(len(k["rows"])), key=lambda i: -int(k["rows"][i][ni]))[:top]
        for i in sorted(order):
            r = k["rows"][i]
            st = sorted(((int(r[c]), h[c][6:]) for c in stall_cols if int(r[c]) > 0), reverse=True)[:3]
            print(f"   #{i:5d} {100.0 * int(r[ni]) / tot_s:5.1f}%  exec {int(r[ei]):>9d}  {r[si].strip()[:70]:70s} "
                  + " ".join(f"{n}:{v}" for v, n in st))


if __name__ == "__main__":
    main()
