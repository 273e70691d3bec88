#!/bin/bash
# Final evidence run of a round: smoke, parity tests, bench (both arms), ncu launch list + full captures.
TAG=${1:-r1final}
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -3 gpurun_out/smoke_$TAG.log
python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/pytest_$TAG.log; tail -2 gpurun_out/pytest_$TAG.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cut -c1-330 gpurun_out/bench_$TAG.json
python scripts/bench_configs.py > gpurun_out/configs_$TAG.json 2> gpurun_out/configs_$TAG.err
grep -A3 "C5_" gpurun_out/configs_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench_$TAG.log 2>&1
for spec in "64 conv3x3_tcp conv64" "128 conv3x3_tcp2 conv128" "64 wgrad3x3_tc wgrad64" "128 wgrad3x3_tc wgrad128"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on -k regex:$2 -c 8 -o gpurun_out/full_$3_$TAG -f \
      python scripts/prof_odeblock.py $1 512 > gpurun_out/ncu_$3_$TAG.log 2>&1
  python scripts/ncu_summary.py gpurun_out/full_$3_$TAG.ncu-rep > gpurun_out/ncu_full_$3_$TAG.txt 2>&1
  ncu -i gpurun_out/full_$3_$TAG.ncu-rep --page source --csv > /tmp/src_$3.csv 2>/dev/null
  python scripts/sass_hot.py /tmp/src_$3.csv 0 > gpurun_out/ncu_sass_$3_$TAG.txt 2>&1
  if [ "$3" != "conv64" ]; then rm -f gpurun_out/full_$3_$TAG.ncu-rep; fi      # gpurun copies back at most 64 MiB
done
ls -la gpurun_out | grep $TAG
