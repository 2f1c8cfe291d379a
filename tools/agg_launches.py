"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys


def load(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    for row in csv.DictReader(lines):
        try:
            rows.append((row['Kernel Name'], float(row['Metric Value'].replace(',', ''))))
        except Exception:
            pass
    return rows


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = load(path)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v in rows:
        k = re.sub(r'\(.*', '', k)
        k = re.sub(r'^void ', '', k).replace('<unnamed>::', '')[:100]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print('%d launches, total %.3f ms (device time, cold-cache, serialised)' % (len(rows), tot / 1e6))
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print('%9.3f ms %5.1f%%  n=%5d avg=%8.1f us  %s' % (t / 1e6, 100 * t / tot, n, t / n / 1e3, k))


if __name__ == '__main__':
    main()
