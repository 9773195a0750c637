"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: share of ours (tamtr::) vs library."""
import collections
import csv
import re
import sys


def main(path, top=45):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"]
        v = v / 1000 if u in ("ns", "nsecond") else v * 1000 if u in ("ms", "msecond") else v
        name = re.sub(r"\(.*", "", d["Kernel Name"])
        name = re.sub(r"<.*", "", name)[:80]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if "tamtr::" in k)
    print(f"{path}: {sum(v[0] for v in agg.values())} launches, {tot:.1f} us total, tamtr:: {ours:.1f} us ({100 * ours / tot:.1f} %)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{v[1]:10.1f} us {v[0]:5d}x  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
