#!/bin/bash
# Peer-memory exchange evidence (run with gpurun --gpus N): the exchange step alone, then the headline and config 4 with the
# peer kernel and with NCCL, then config 5 with the single count reduction.
N=${1:-2}
TAG=${2:-r2peer}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
timeout 300 $RUN scripts/bench_exchange.py > gpurun_out/exchange_${TAG}_n$N.json 2> gpurun_out/exchange_${TAG}_n$N.err; cat gpurun_out/exchange_${TAG}_n$N.json; tail -3 gpurun_out/exchange_${TAG}_n$N.err
timeout 300 $RUN bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
python - gpurun_out/bench_${TAG}_n$N.json <<'P'
import json, sys
d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
print(d["n_gpus"], round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"], d["config"]["allreduce"][:160])
P
timeout 300 $RUN bench.py --gpus $N --steps 10 --warmup 3 --nccl-allreduce > gpurun_out/bench_${TAG}_nccl_n$N.json 2> gpurun_out/bench_${TAG}_nccl_n$N.err
python - gpurun_out/bench_${TAG}_nccl_n$N.json <<'P'
import json, sys
d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
print(d["n_gpus"], round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"], d["config"]["allreduce"][:160])
P
timeout 300 $RUN scripts/bench_c4.py > gpurun_out/c4_${TAG}_n$N.json 2> gpurun_out/c4_${TAG}_n$N.err; cat gpurun_out/c4_${TAG}_n$N.json; tail -2 gpurun_out/c4_${TAG}_n$N.err
timeout 300 $RUN scripts/bench_c4.py --nccl-allreduce > gpurun_out/c4_${TAG}_nccl_n$N.json 2> gpurun_out/c4_${TAG}_nccl_n$N.err; cat gpurun_out/c4_${TAG}_nccl_n$N.json
if [ "${3:-c5}" = "c5" ]; then
timeout 400 $RUN scripts/eval_pgd_sweep.py --n-images 8192 --batch 512 --golden tests/golden/pgd_sweep.npz > gpurun_out/c5_${TAG}_n$N.json 2> gpurun_out/c5_${TAG}_n$N.err
grep '^{' gpurun_out/c5_${TAG}_n$N.json | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['images_per_s'],d['seconds'],d['vs_reference_golden'])"
fi
