#!/bin/bash
# light ncu pass over the C = 64 weight gradient in both geometries (option wgrad_htaps 0 / 1)
M="gpu__time_duration.sum,sm__cycles_elapsed.avg,sm__inst_executed_pipe_tensor.sum,l1tex__data_pipe_tc_wavefronts_mem_shared.sum,lts__t_bytes.sum,lts__t_sectors_srcunit_tex.sum,dram__bytes_read.sum,smsp__cycles_active.avg,l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"
for HT in 0 1; do
  MSB_WGRAD_HTAPS=$HT PROF_GRAD=1 ncu --metrics $M --clock-control none -k regex:"wgrad3x3_tc" -s 2 -c 2 --csv \
      --log-file gpurun_out/ncu_wgrad_ht$HT.csv python scripts/prof_conv.py 2 0 512 64 > gpurun_out/ncu_wgrad_ht$HT.log 2>&1
done
python - <<'PY'
import csv, io
for ht in (0, 1):
    f = "gpurun_out/ncu_wgrad_ht%d.csv" % ht
    lines = [l for l in open(f).read().splitlines() if l.startswith('"')]
    rows = list(csv.reader(io.StringIO("\n".join(lines)))); h = rows[0]
    im, iv, iid = h.index('Metric Name'), h.index('Metric Value'), h.index('ID')
    d = {}
    for r in rows[1:]:
        d.setdefault(r[iid], {})[r[im]] = r[iv]
    for k, m in d.items():
        print("htaps=%d launch %s: " % (ht, k) + "  ".join("%s=%s" % (a.split("__")[-1][:44], b) for a, b in m.items()))
PY
