"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel."""
import collections
import csv
import io
import sys


def main(path, top=25):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    tot, cnt = collections.Counter(), collections.Counter()
    for row in csv.DictReader(io.StringIO("".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        name = row["Kernel Name"][:80]
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"total {T:.1f} us over {sum(cnt.values())} launches (cold-cache, serialised: compare shares)")
    for n, v in tot.most_common(top):
        print(f"{v:10.1f} us {100 * v / T:5.1f}%  x{cnt[n]:4d}  avg {v / cnt[n]:8.1f} us  {n}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
