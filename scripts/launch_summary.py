"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total ms and share per kernel."""
import csv, sys, collections, re
path = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
h = rows[0]
ik, iv, im, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    v = v / 1e6 if r[iu] in ("ns", "nsecond") else (v / 1e3 if r[iu] in ("us", "usecond") else v)
    name = re.sub(r"\(.*", "", r[ik])
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("# %s" % " ".join(sys.argv[2:]))
print("# per-launch times are cold-cache/serialised: compare SHARES with bench.py's roofline.conv_share_of_step / wgrad_share_of_step, not absolutes")
print("%-72s %6s %10s %6s" % ("kernel", "n", "ms", "share"))
for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-72s %6d %10.3f %5.1f%%" % (name[-72:], n, ms, 100 * ms / tot))
conv = sum(ms for k, (n, ms) in agg.items() if "conv3x3_tc" in k)
wg = sum(ms for k, (n, ms) in agg.items() if "wgrad3x3_tc" in k)
other = sum(ms for k, (n, ms) in agg.items() if "msb" not in k)
print("# conv3x3_tct + conv3x3_tcp2 (+ tcp) share %.1f%%, wgrad3x3_tc share %.1f%%, non-msb kernels share %.1f%%" % (100 * conv / tot, 100 * wg / tot, 100 * other / tot))
