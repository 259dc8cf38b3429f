"""Per-source-line / per-phase instruction budget of one kernel from an ncu report captured with --import-source on.

usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:NAME > src.csv
       python profiles/phase_budget.py src.csv [phases.json]
phases.json: {"file.cuh": [[first_line, last_line, "phase name"], ...]}
"""
import collections
import csv
import json
import sys


def load(path):
    cur, ie, ns, agg = None, None, None, collections.OrderedDict()
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            ie, ns = r.index("Instructions Executed"), r.index("# Samples")
        elif r[0] not in ("", "Function Name") and ie is not None:
            try:
                agg[(cur, int(r[0]))] = (int(r[ie]), int(r[ns]), r[1][:100])
            except ValueError:
                pass
    return agg


def main():
    agg = load(sys.argv[1])
    phases = json.load(open(sys.argv[2])) if len(sys.argv) > 2 else {}
    tot = sum(v[0] for v in agg.values())
    ts = sum(v[1] for v in agg.values()) or 1
    print(f"total warp instructions {tot/1e6:.2f} M, stall samples {ts}")
    if phases:
        ph = collections.OrderedDict()
        for (f, l), (i, s, _) in agg.items():
            name = f + ": other"
            for a, b, n in phases.get(f, []):
                if a <= l <= b:
                    name = n
                    break
            x = ph.setdefault(name, [0, 0])
            x[0] += i
            x[1] += s
        print("| phase | warp instructions (M) | % of instructions | % of stall samples (~cycles) |\n|---|---|---|---|")
        for n, (i, s) in sorted(ph.items(), key=lambda kv: -kv[1][0]):
            print(f"| {n} | {i/1e6:.2f} | {100*i/tot:.1f} | {100*s/ts:.1f} |")
    else:
        for (f, l), (i, s, src) in sorted(agg.items()):
            if i > tot * 0.004:
                print(f"{f}:{l:5d} {i/1e6:8.2f}M {100*i/tot:5.1f}% samp {100*s/ts:5.1f}%  {src}")


if __name__ == "__main__":
    main()
