"""Hot SASS lines of one kernel from `ncu --page source --csv`: sample share, executions, dominant stall reasons.

    ncu -i REP --page source --csv --kernel-name regex:NAME > src.csv ; python tools/ncu_hot.py src.csv [min_pct] [lo hi]
"""
import csv
import sys


def main(path, min_pct=0.4, lo=None, hi=None):
    rows = list(csv.reader(open(path)))
    sections, cur = [], None                         # one section per kernel: "Kernel Name" row, header row, data rows
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = [r[1], None, []]
            sections.append(cur)
        elif cur is not None and cur[1] is None:
            cur[1] = r
        elif cur is not None:
            cur[2].append(r)
    want = sys.argv[5] if len(sys.argv) > 5 else ""
    name, hdr, data = next(sec for sec in sections if want in sec[0])
    print(name[:100])
    data = [r for r in data if len(r) == len(hdr)]
    ix = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {c: sum(int(r[ix[c]]) for r in data) for c in cols}
    print("samples", tot, "instructions", len(data))
    print({c[6:]: round(100 * v / tot, 1) for c, v in agg.items() if v > tot * 0.005})
    for i, r in enumerate(data):
        s = int(r[ix["# Samples"]])
        if (hi and lo <= i < hi) or (not hi and s > tot * min_pct / 100):
            why = " ".join(f"{c[6:]}={r[ix[c]]}" for c in cols if int(r[ix[c]]) > max(1, s * 0.25))
            print(f"{i:5d} {r[ix['Source']].strip()[:64]:64s} {s:6d} x{r[ix['Instructions Executed']]:>10s} {why}")


if __name__ == "__main__":
    a = sys.argv
    main(a[1], float(a[2]) if len(a) > 2 else 0.4, int(a[3]) if len(a) > 4 else None, int(a[4]) if len(a) > 4 else None)
