"""Turn an `ncu --metrics gpu__time_duration.sum --csv` log into the per-kernel share table kept under profiles/.
Usage: python tools/ncu_launch_table.py launches.csv "title" "command" > profiles/xxx.md"""
import csv
import sys
from collections import defaultdict


def main():
    path, title, command = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    iname, imetric, ival, iunit = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= ival or r[imetric] != "gpu__time_duration.sum":
            continue
        v = float(r[ival].replace(",", ""))
        unit = r[iunit]
        us = v / 1e3 if unit.startswith("ns") else (v * 1e3 if unit.startswith("ms") else v)
        rows.append((r[iname], us))
    agg = defaultdict(lambda: [0, 0.0])
    for name, us in rows:
        agg[name][0] += 1
        agg[name][1] += us
    total = sum(v[1] for v in agg.values())
    print("# %s\n" % title)
    print("Command (under gpurun): `%s`\n" % command)
    print("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes. %d launches captured.\n" % len(rows))
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f%% |" % (name[:110], cnt, us, 100 * us / total))
    print("\nTotal %.1f us over %d launches." % (total, len(rows)))


if __name__ == "__main__":
    main()
