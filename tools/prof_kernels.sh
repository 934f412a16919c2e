# ncu --set full captures of the whole-job launches of the FIR kernels (one capture per kernel), after a plain run
export WHOLE=1 S=256 NB=47 REP=2
python tools/profile_run.py || exit 1
for k in k_frontend_stream k_twin_bpf k_audio_u1; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_r1_$k python tools/profile_run.py > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
