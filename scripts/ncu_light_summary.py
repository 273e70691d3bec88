"""Summarise the light counter pass of scripts/gpu_ncu_r2.sh (ncu --metrics ... --csv): per launch duration, elapsed
cycles, the non-sampled tensor-pipe counter (sm__pipe_tensor_subpipe_hmma_cycles_active_realtime, summed over the 4
sub-partitions of an SM -> / 4) as a fraction of the elapsed cycles, and DRAM bytes.
    python scripts/ncu_light_summary.py gpurun_out/ncu_light_c64_r2.csv"""
import csv, io, sys
for f in sys.argv[1:]:
    lines = [l for l in open(f).read().splitlines() if l.startswith('"')]
    rows = list(csv.reader(io.StringIO("\n".join(lines))))
    h = rows[0]; ik = h.index('Kernel Name'); im = h.index('Metric Name'); iv = h.index('Metric Value'); iid = h.index('ID')
    d = {}
    def num(s):
        try: return float(s.replace(',', ''))
        except ValueError: return float('nan')
    for r in rows[1:]:
        d.setdefault((int(r[iid]), r[ik][:64]), {})[r[im]] = num(r[iv])
    print("#", f)
    print("# util(inst) = sm__inst_executed_pipe_tensor.sum x clocks per MMA / issuing SMs / elapsed cycles: conv3x3_tct 32 clk (N = 64, A in TMEM)"
          " x 148 SMs; wgrad 96 clk (N = 192) x 148; conv3x3_tcp2 96 clk average (N = 256 + N = 128 per pair) x 74 pairs")
    for (i, k), m in sorted(d.items()):
        el = m['sm__cycles_elapsed.avg']; hm = m['sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg']
        ti = m.get('sm__inst_executed_pipe_tensor.sum', float('nan'))
        clk, nsm = (32, 148) if 'conv3x3_tct' in k else ((96, 74) if 'tcp2' in k else (96, 148))
        print("%-58s util(inst) %4.1f %%" % (k[:58], 100.0 * ti * clk / nsm / el), end="  |  ")
        print("%-64s %7.1f us  elapsed %7.0f clk @%.2f GHz  hmma_active/4 %7.0f = %4.1f %% of elapsed  tensor inst %8.0f  dram %4.0f MB  (sampled pct metric %4.1f)"
              % (k, m['gpu__time_duration.sum'] / 1e3, el, m['sm__cycles_elapsed.avg.per_second'] / 1e9, hm / 4, 100 * hm / 4 / el,
                 m.get('sm__inst_executed_pipe_tensor.sum', float('nan')),
                 (m['dram__bytes_read.sum'] + m['dram__bytes_write.sum']) / 1e6,
                 m['sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed']))
