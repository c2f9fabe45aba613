"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`).
    python tools/launch_list_summary.py gpurun_out/dubo_launches.csv [top_n]"""
import collections
import csv
import sys

path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
lines = [l for l in open(path) if not l.startswith("==")]
tot, cnt = collections.OrderedDict(), collections.Counter()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"][:78]
    v, unit = float(row["Metric Value"].replace(",", "")), row["Metric Unit"]
    ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
    tot[name] = tot.get(name, 0.0) + ms
    cnt[name] += 1
T = sum(tot.values())
print(f"total {T:.3f} ms over {sum(cnt.values())} launches (cold-cache, serialised: shares, not absolutes)")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{v:8.3f} ms {100 * v / T:5.1f}%  x{cnt[k]:3d}  {k}")
