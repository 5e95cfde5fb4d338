"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel and grid (development aid).
   python tools/launch_summary.py <launches.csv> [first-kernel-substring last-kernel-substring [occurrence]]
With the two substrings only the launches of ONE pass between them (the given occurrence, default the last) are summed."""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    start = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[start]
    ki, gi, vi = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value")
    names = []
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        n = re.sub(r"\(.*", "", r[ki]).replace("cddpm::<unnamed>::", "").replace("void ", "")
        names.append((n, r[gi], float(r[vi].replace(",", "")) / 1e3))
    if len(sys.argv) >= 4:
        first, last = sys.argv[2], sys.argv[3]
        occ = int(sys.argv[4]) if len(sys.argv) > 4 else -1
        begins = [i for i, n in enumerate(names) if first in n[0]]
        i0 = begins[occ]
        j = i0
        while last not in names[j][0]:
            j += 1
        names = names[i0:j + 1]
    tot = sum(t for _, _, t in names)
    agg = collections.OrderedDict()
    for n, g, t in names:
        a = agg.setdefault((n, g), [0, 0.0])
        a[0] += 1
        a[1] += t
    print(f"total {tot:.1f} us over {len(names)} launches")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:9.1f} us {100 * t / tot:5.1f}%  x{c:3d}  avg {t / c:7.1f}  {k[0][:70]} grid {k[1]}")


if __name__ == "__main__":
    main()
