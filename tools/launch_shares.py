"""Per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in rows:
    name = r[4].split("(")[0].replace("void ", "")
    tot[name] += float(r[-1])
    cnt[name] += 1
total = sum(tot.values())
print(f"{len(rows)} launches, {total / 1e3:.1f} us in total (cold-cache, serialised: compare SHARES)")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k:28s} n={cnt[k]:4d} avg={v / cnt[k] / 1e3:9.2f} us share={100 * v / total:5.1f}%")
