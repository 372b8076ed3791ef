"""Summarise an .ncu-rep (read on the CPU box): key throughput metrics per captured launch and, with --source,
the hottest source lines by sampled stalls.   usage: python profiles/ncu_summary.py REP [--source N]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
STALL = "smsp__average_warps_issue_stalled_"


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    for r in data:
        print("==", r[name_i])
        for i, h in enumerate(hdr):
            if h in KEYS:
                print(f"   {h:75s} {r[i]:>16s} {units[i]}")
        st = sorted(((float(r[i]), h[len(STALL):-len('_per_issue_active.ratio')]) for i, h in enumerate(hdr)
                     if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i]), reverse=True)
        print("   stalls/issue:", ", ".join(f"{n} {v:.2f}" for v, n in st[:8]))


def source(rep, top):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
    blocks = out.split("\n\n")
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    acc = []
    for r in rows:
        if "Source" in r and any("Sampl" in c for c in r):
            hdr = r
            si = hdr.index("Source")
            ci = [i for i, c in enumerate(hdr) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)"]
            ci = ci[0] if ci else None
            li = hdr.index("#") if "#" in hdr else 0
            if acc:
                show(acc, top)
            acc = []
            continue
        if hdr and ci is not None and len(r) > max(si, ci):
            try:
                acc.append((int(r[ci] or 0), r[li], r[si].strip()))
            except ValueError:
                pass
    if acc:
        show(acc, top)


def show(acc, top):
    tot = sum(a[0] for a in acc) or 1
    print(f"-- kernel view, {tot} samples")
    for n, ln, src in sorted(acc, reverse=True)[:top]:
        print(f"   {100.0 * n / tot:5.1f}%  L{ln:>4s}  {src[:150]}")


if __name__ == "__main__":
    rep = sys.argv[1]
    raw(rep)
    if "--source" in sys.argv:
        source(rep, int(sys.argv[sys.argv.index("--source") + 1]))
