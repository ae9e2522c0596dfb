"""Summarise one SAC learn out of an ncu launch list of scripts/sac_launches.py (gpu__time_duration.sum per launch):
    python scripts/sac_launch_summary.py gpurun_out/launches_sac.csv "header text" > profiles/launches_rNN_sac_learn.txt
One learn = the launches between two consecutive k_polyak_pack launches of the eager warm-up loop."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
ix = {h: i for i, h in enumerate(rows[hdr])}
recs = []
for r in rows[hdr + 1:]:
    try:
        recs.append((r[ix["Kernel Name"]], float(r[ix["Metric Value"]]) / 1000.0, r[ix["Grid Size"]]))
    except Exception:
        pass
idx = [i for i, (k, _, _) in enumerate(recs) if "k_polyak_pack" in k]
one = recs[idx[1] + 1: idx[2] + 1]
# the learner's own kernels only (the eager loop also launches torch's randn / copy kernels between learns when eps is not given)
own = [(k, v, g) for k, v, g in one if "b2q" in k or "k_" in k.split("(")[0]]
agg = collections.OrderedDict()
for k, v, g in own:
    name = k.split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    a = agg.setdefault(name, [0, 0.0, set()])
    a[0] += 1; a[1] += v; a[2].add(g)
print("%s; sum %.1f us over %d launches" % (sys.argv[2] if len(sys.argv) > 2 else "one SAC learn", sum(v for _, v, _ in own), len(own)))
for name, (n, tot, grids) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-46s n=%3d avg=%6.1f us tot=%7.1f us grids=%s" % (name[-46:], n, tot / n, tot, sorted(grids)))
print("launch order:")
for k, v, g in own:
    print("  %-40s %6.1f us  %s" % (k.split("(")[0].replace("void ", "").replace("<unnamed>::", "")[-40:], v, g))
