import collections, csv, sys
src, dst = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
lines = [l for l in open(src) if not l.startswith("==")]
order = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", "")); unit = row["Metric Unit"]
    us = v / 1000 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000)
    order.append((row["Kernel Name"][:90], us))
tot = sum(u for _, u in order)
print("kernels in step: %d, serialized total %.1f us" % (len(order), tot))
agg = collections.defaultdict(lambda: [0, 0.0])
for n, u in order:
    agg[n][0] += 1; agg[n][1] += u
for n, (c, u) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print("%8.1f us %5.1f%%  x%-3d %s" % (u, 100 * u / tot, c, n))
if dst:
    open(dst, "w").write("\n".join("%9.1f  %s" % (u, n) for n, u in order))
