#!/bin/bash
# Final evidence run of round 2: smoke, parity tests, bench (both arms), per-config numbers, ncu launch list.
TAG=${1:-r2final}
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -3 gpurun_out/smoke_$TAG.log
python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/pytest_$TAG.log; tail -2 gpurun_out/pytest_$TAG.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cut -c1-330 gpurun_out/bench_$TAG.json
python scripts/bench_configs.py > gpurun_out/configs_$TAG.json 2> gpurun_out/configs_$TAG.err
grep -A4 "C1_\|C5_" gpurun_out/configs_$TAG.json | head -24
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench_$TAG.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_$TAG.csv > gpurun_out/launches_${TAG}_summary.txt 2>&1; head -30 gpurun_out/launches_${TAG}_summary.txt
ls -la gpurun_out | grep $TAG
