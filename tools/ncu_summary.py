"""Summarise `ncu --csv` output (raw page or --metrics log) per kernel and update profiles/ncu_traffic.json.

    python tools/ncu_summary.py <csv> --label r2_config2 [--traffic-key "skinny_tma_kernel|m=37032|n=6750|k=10|gpus=1" --match skinny_tma_kernel]

Prints, per kernel name, the number of launches and the mean of every numeric metric; with --traffic-key stores
mean(dram__bytes_read.sum + dram__bytes_write.sum) per launch of the kernels matching --match, which bench.py reports
as `roofline.traffic`."""
import argparse
import csv
import json
import os
import re
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6,
              "msecond": 1e6, "nsecond": 1.0, "second": 1e9}


def read(path):
    """Long format (one row per launch x metric: columns 'Kernel Name','Metric Name','Metric Unit','Metric Value')
    or wide format (raw page: one row per launch, one column per metric, second line = units)."""
    lines = [l for l in open(path, errors="replace") if not l.startswith("==")]
    rows = list(csv.reader(lines))
    rows = [r for r in rows if r]
    head = rows[0]
    out = defaultdict(lambda: defaultdict(list))          # kernel -> metric -> values per launch
    if "Metric Name" in head:
        ik, im, iu, iv, iid = (head.index("Kernel Name"), head.index("Metric Name"), head.index("Metric Unit"),
                               head.index("Metric Value"), head.index("ID"))
        for r in rows[1:]:
            try:
                v = float(r[iv].replace(",", ""))
            except ValueError:
                continue
            out[r[ik]][r[im] + " [" + r[iu] + "]"].append(v)
    else:
        units = rows[1]
        ik = head.index("Kernel Name")
        for r in rows[2:]:
            for j, name in enumerate(head):
                if j == ik or j >= len(r):
                    continue
                try:
                    v = float(r[j].replace(",", ""))
                except ValueError:
                    continue
                out[r[ik]][name + " [" + units[j] + "]"].append(v)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--label", default="")
    ap.add_argument("--traffic-key", default=None)
    ap.add_argument("--match", default=None)
    ap.add_argument("--only", default=None, help="regex of metric names to print")
    a = ap.parse_args()
    data = read(a.csv)
    for kern, metrics in data.items():
        n = max(len(v) for v in metrics.values())
        print("%s\n  launches: %d" % (kern[:150], n))
        for name in sorted(metrics):
            if a.only and not re.search(a.only, name):
                continue
            v = metrics[name]
            print("  %-78s mean %.6g  min %.6g  max %.6g" % (name, sum(v) / len(v), min(v), max(v)))
    if a.traffic_key:
        tot, cnt = 0.0, 0
        for kern, metrics in data.items():
            if a.match and a.match not in kern:
                continue
            rd = [(name, v) for name, v in metrics.items() if name.startswith("dram__bytes_read.sum ")]
            wr = [(name, v) for name, v in metrics.items() if name.startswith("dram__bytes_write.sum ")]
            if not rd or not wr:
                continue
            def to_bytes(name, vals):
                unit = name[name.index("[") + 1:-1]
                return [x * UNIT_SCALE.get(unit, 1.0) for x in vals]
            r, w = to_bytes(*rd[0]), to_bytes(*wr[0])
            tot += sum(r) + sum(w)
            cnt += len(r)
        if cnt:
            path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
            table = json.load(open(path)) if os.path.exists(path) else {}
            table[a.traffic_key] = {"dram_bytes_per_launch": tot / cnt, "launches": cnt,
                                    "source": "profiles/%s (dram__bytes_read.sum + dram__bytes_write.sum, mean per launch)" % a.label}
            json.dump(table, open(path, "w"), indent=1, sort_keys=True)
            print("traffic[%s] = %.6g bytes per launch over %d launches" % (a.traffic_key, tot / cnt, cnt))
        else:
            print("no dram__bytes metrics for --match %r" % a.match, file=sys.stderr)


if __name__ == "__main__":
    main()
