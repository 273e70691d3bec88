#!/bin/bash
# Round-2 final ncu evidence: full captures of both convolution kernels of the final build (conv3x3_tct for C = 64,
# conv3x3_tcp2 for C = 128) and the light non-sampled tensor-pipe counter pass over every tcgen05 kernel of a
# forward + backward RK2 step at B = 512.
TAG=${1:-r2final}
M="gpu__time_duration.sum,sm__cycles_elapsed.avg,sm__cycles_elapsed.avg.per_second,sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,l1tex__data_pipe_tc_wavefronts_mem_shared.sum,lts__t_bytes.sum"
python scripts/prof_conv.py 2 0 512 64 > gpurun_out/p_full_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv3x3_tct -s 4 -c 4 -f -o gpurun_out/full_tct_$TAG \
    python scripts/prof_conv.py 2 0 512 64 > gpurun_out/ncu_full_tct_$TAG.log 2>&1
python scripts/prof_conv.py 2 0 512 128 > gpurun_out/p_full128_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv3x3_tcp2 -s 4 -c 4 -f -o gpurun_out/full_tcp2_$TAG \
    python scripts/prof_conv.py 2 0 512 128 > gpurun_out/ncu_full_tcp2_$TAG.log 2>&1
for C in 64 128; do
  PROF_GRAD=1 python scripts/prof_conv.py 2 0 512 $C > gpurun_out/p_light_$TAG.log 2>&1 && \
  PROF_GRAD=1 ncu --metrics $M --clock-control none -k regex:"conv3x3_t|wgrad3x3_tc" -s 12 -c 12 --csv \
      --log-file gpurun_out/ncu_light_c${C}_$TAG.csv python scripts/prof_conv.py 2 0 512 $C > gpurun_out/ncu_light_$TAG.log 2>&1
done
ls -la gpurun_out | grep $TAG
