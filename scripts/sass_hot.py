"""Hot SASS regions / opcode mix of ONE launch of an ncu source-page CSV (ncu -i rep --page source --csv)."""
import csv, sys, collections
path, which = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(open(path)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
starts.append(len(rows))
blk = rows[starts[which]:starts[which + 1]]
print(blk[0][1][:100])
h = blk[1]; ia = h.index('Source'); ie = h.index('Instructions Executed'); iss = h.index('# Samples'); it = h.index('Avg. Threads Executed')
body = [r for r in blk[2:] if len(r) > iss and r[ie].isdigit()]
tot_s = sum(int(r[iss] or 0) for r in body); tot_i = sum(int(r[ie]) for r in body)
print('total samples', tot_s, 'total warp instr', tot_i)
agg = collections.Counter()
for r in body:
    op = r[ia].strip().split()
    o = op[1] if op[0].startswith('@') else op[0]
    agg['.'.join(o.split('.')[:2]) if o.startswith(('LDG', 'STG', 'LDTM', 'SHFL', 'MUFU', 'SYNCS')) else o.split('.')[0]] += int(r[ie])
print(' '.join('%s:%.1f%%' % (o, 100 * n / tot_i) for o, n in agg.most_common(24)))
prev = None; out = []
for i, r in enumerate(body):
    key = (int(r[ie]), r[it])
    if key != prev:
        out.append([i, i, key, 0, 0]); prev = key
    out[-1][1] = i; out[-1][3] += 1; out[-1][4] += int(r[iss] or 0)
for s, e, k, c, sm in out:
    if sm > tot_s * 0.01 or k[0] * c > tot_i * 0.01:
        print("rows %5d-%5d exec/instr %9d thr %5s ninstr %4d warp-instr %5.1f%% samples %5.1f%%  first: %s" % (
            s, e, k[0], k[1], c, 100.0 * k[0] * c / tot_i, 100.0 * sm / tot_s, body[s][ia].strip()[:50]))
