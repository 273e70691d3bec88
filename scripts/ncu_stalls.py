"""Per-SASS-line stall breakdown of one launch of an ncu source-page CSV (ncu -i rep --page source --csv): the hottest
lines by warp-stall samples with their dominant stall reasons, plus the totals by reason.
    python scripts/ncu_stalls.py src.csv [launch index] [top N]"""
import csv, sys, collections
path = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
blk = rows[starts[which]:starts[which + 1]]
print(blk[0][1][:120])
h = blk[1]
ia = h.index('Source'); iss = h.index('# Samples'); ie = h.index('Instructions Executed')
stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith('stall_') or c.lower().startswith('warp stall') or 'Stall' in c]
body = [r for r in blk[2:] if len(r) > iss and r[ie].isdigit()]
tot = sum(int(r[iss] or 0) for r in body)
print("total samples", tot, "| stall columns:", [c for _, c in stall_cols][:40])
agg = collections.Counter()
for r in body:
    for i, c in stall_cols:
        try: agg[c] += int(r[i] or 0)
        except ValueError: pass
print("by reason:", ", ".join("%s=%.1f%%" % (c, 100.0 * n / max(tot, 1)) for c, n in agg.most_common(14)))
order = sorted(range(len(body)), key=lambda k: -int(body[k][iss] or 0))[:top]
for k in sorted(order):
    r = body[k]
    rs = sorted(((int(r[i] or 0), c) for i, c in stall_cols if (r[i] or '0').isdigit()), reverse=True)[:3]
    print("%5d %5.1f%% exec %8s  %-70s %s" % (k, 100.0 * int(r[iss] or 0) / max(tot, 1), r[ie], r[ia].strip()[:70],
                                              " ".join("%s:%d" % (c.replace('stall_', ''), n) for n, c in rs if n)))
