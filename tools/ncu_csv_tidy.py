"""Reduces raw `ncu --csv --log-file` outputs to the small tables kept under profiles/.

  python tools/ncu_csv_tidy.py launches <raw.csv> <out.csv>   per kernel (name + grid): launches, total us, share
  python tools/ncu_csv_tidy.py graph    <raw.csv> <out.csv>   per profiled graph launch: metric, unit, value
"""
import collections
import csv
import sys


def rows(path):
    lines = [ln for ln in open(path) if ln.startswith('"')]
    return list(csv.DictReader(lines))


def main():
    mode, src, dst = sys.argv[1:4]
    rs = rows(src)
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        if mode == "launches":
            tot, cnt = collections.Counter(), collections.Counter()
            for r in rs:
                if r["Metric Name"] != "gpu__time_duration.sum":
                    continue
                key = "%s grid%s" % (r["Kernel Name"].split("(")[0][:70], r["Grid Size"])
                tot[key] += float(r["Metric Value"].replace(",", "")) / 1e3
                cnt[key] += 1
            total = sum(tot.values())
            w.writerow(["kernel", "launches", "total_us", "share_pct"])
            for k, v in tot.most_common():
                w.writerow([k, cnt[k], "%.1f" % v, "%.2f" % (100.0 * v / total)])
        else:
            w.writerow(["graph_launch", "metric", "unit", "value"])
            for r in rs:
                w.writerow([r["ID"], r["Metric Name"], r["Metric Unit"], r["Metric Value"].replace(",", "")])


if __name__ == "__main__":
    main()
