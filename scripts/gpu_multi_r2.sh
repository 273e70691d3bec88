#!/bin/bash
# Multi-GPU evidence of round 2 (run with gpurun --gpus N): weak / strong scaling of the headline, C4 and C5 at N ranks.
N=${1:-2}
TAG=${2:-r2}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$RUN bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
cut -c1-200 gpurun_out/bench_${TAG}_n$N.json
$RUN bench.py --gpus $N --steps 10 --warmup 3 --batch $((512 / N)) > gpurun_out/bench_strong_${TAG}_n$N.json 2> gpurun_out/bench_strong_${TAG}_n$N.err
cut -c1-200 gpurun_out/bench_strong_${TAG}_n$N.json
$RUN scripts/bench_c4.py > gpurun_out/c4_${TAG}_n$N.json 2> gpurun_out/c4_${TAG}_n$N.err
cat gpurun_out/c4_${TAG}_n$N.json
$RUN scripts/eval_pgd_sweep.py --n-images 8192 --batch 512 --golden tests/golden/pgd_sweep.npz > gpurun_out/c5_${TAG}_n$N.json 2> gpurun_out/c5_${TAG}_n$N.err
cut -c1-700 gpurun_out/c5_${TAG}_n$N.json
