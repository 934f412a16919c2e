#!/bin/bash
# Round-2 evidence in one GPU call: launch list + DRAM bytes of one bench step, and `ncu --set full` captures of the hot kernels.
set -x
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_bench.log 2>&1 || exit 1
# every library kernel of the run: duration and DRAM bytes; tools/step_traffic.py picks the two timed steps (calls 3 and 4) out of the list
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_ -c 3000 --csv \
    --log-file gpurun_out/r2_step_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_step.log 2>&1
python tools/step_traffic.py gpurun_out/r2_step_launches.csv 3 2 > gpurun_out/r2_step_traffic_table.md
# whole-job launches of the FIR kernels (one sub-chunk per call)
export WHOLE=1 S=256 NB=48 REP=2 MODE=0
python tools/profile_run.py > gpurun_out/plain_whole.log 2>&1 || exit 1
for k in k_frontend_stream k_bpf_mixed k_audio_u1 k_pll_table_ops k_pll_predict k_nco_phase; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_r2_$k python tools/profile_run.py > gpurun_out/ncu_$k.log 2>&1
done
# the serial loop inside a pipelined call: 8-block sub-chunk of the second pass (periodic input: the second pass continues the first)
export WHOLE=0 NB=24
python tools/profile_run.py > gpurun_out/plain_pipe.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_pll_sel -s 9 -c 1 -f -o gpurun_out/prof_r2_k_pll_sel python tools/profile_run.py > gpurun_out/ncu_k_pll_sel.log 2>&1
# polyphase audio kernel, mode 2 (1024 streams x 12 blocks, whole-job launch)
export WHOLE=1 S=1024 NB=12 MODE=2
python tools/profile_run.py > gpurun_out/plain_m2.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_audio_poly -s 1 -c 1 -f -o gpurun_out/prof_r2_k_audio_poly_ts python tools/profile_run.py > gpurun_out/ncu_poly.log 2>&1
# one RDS kernel: the 19/120 resampler (4096 streams x 6 blocks)
export WHOLE=0 S=4096 NB=6 MODE=0 RDS=1
python tools/profile_run.py > gpurun_out/plain_rds.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_rds_resample -s 1 -c 1 -f -o gpurun_out/prof_r2_k_rds_resample python tools/profile_run.py > gpurun_out/ncu_rds.log 2>&1
ls -la gpurun_out/*.ncu-rep
