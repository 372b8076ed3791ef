"""Per-source-line attribution of an .ncu-rep for one kernel: joins ncu's per-SASS-instruction samples / executed
counts with nvdisasm's line table of the SAME build (cubin extracted from the .so).
usage: python profiles/ncu_lines.py REP SO kernel-substring [top]"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def line_table(so, want):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, capture_output=True)
    for f in os.listdir(d):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-c", "-g", os.path.join(d, f)], capture_output=True, text=True).stdout
        if want not in txt:
            continue
        lines, cur, active, out = txt.split("\n"), None, False, []
        for ln in lines:
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                active = want in m.group(1)
                continue
            if not active:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", ln):
                out.append(cur)
        if out:
            return out
    return []


def main():
    rep, so, want = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    table = line_table(so, want)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur, ks = None, []
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": [], "hdr": None}
            ks.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur and cur["hdr"] and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    for k in ks:
        if want not in k["name"]:
            continue
        h = k["hdr"]
        ni, ei = h.index("# Samples"), h.index("Instructions Executed")
        stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        if len(table) != len(k["rows"]):
            print(f"warning: {len(table)} disassembled vs {len(k['rows'])} profiled instructions (different build?)")
        agg = defaultdict(lambda: [0, 0, defaultdict(int)])
        for i, r in enumerate(k["rows"]):
            key = table[i] if i < len(table) else None
            a = agg[key]
            a[0] += int(r[ni]); a[1] += int(r[ei])
            for c in stall_cols:
                a[2][h[c][6:]] += int(r[c])
        tot_s = sum(a[0] for a in agg.values()) or 1
        tot_e = sum(a[1] for a in agg.values()) or 1
        src = {}
        print(f"== {k['name']}: {tot_e} warp-instr, {tot_s} samples")
        for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
            text = ""
            if key:
                path = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", key[0])
                if key[0] not in src and os.path.exists(path):
                    src[key[0]] = open(path).read().split("\n")
                if key[0] in src and key[1] - 1 < len(src[key[0]]):
                    text = src[key[0]][key[1] - 1].strip()[:90]
            st = sorted(a[2].items(), key=lambda kv: -kv[1])[:3]
            print(f"  {100.0 * a[0] / tot_s:5.1f}% smp {100.0 * a[1] / tot_e:5.1f}% exe  {str(key):28s} {text:90s} " +
                  " ".join(f"{n}:{v}" for n, v in st))
        break


if __name__ == "__main__":
    main()
