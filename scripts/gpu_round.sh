#!/bin/bash
# One GPU session: parity tests, bench, torch-profiler breakdown, ncu launch list, ncu full captures.
set -x
TAG=${1:-r1g}
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log
cat gpurun_out/pytest_$TAG.log | tail -5
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 3000 gpurun_out/bench_$TAG.json
python scripts/profile_step.py 512 > gpurun_out/profstep_$TAG.log 2>&1
head -40 gpurun_out/profstep_$TAG.log
if [ "$2" = "ncu" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_$TAG.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -c 8 -o gpurun_out/full_conv64_$TAG -f \
      python scripts/prof_odeblock.py 64 512 > gpurun_out/ncu_c64_$TAG.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -c 8 -o gpurun_out/full_conv128_$TAG -f \
      python scripts/prof_odeblock.py 128 512 > gpurun_out/ncu_c128_$TAG.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:wgrad3x3_tc -c 4 -o gpurun_out/full_wgrad64_$TAG -f \
      python scripts/prof_odeblock.py 64 512 > gpurun_out/ncu_w64_$TAG.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:wgrad3x3_tc -c 4 -o gpurun_out/full_wgrad128_$TAG -f \
      python scripts/prof_odeblock.py 128 512 > gpurun_out/ncu_w128_$TAG.log 2>&1
fi
ls -la gpurun_out | tail -20
