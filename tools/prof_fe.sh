# ncu full capture of one whole-job front-end launch (variant in $1), after a plain run of the same command
set -x
V=${1:-stream}
[ -n "$UBENCH" ] && ./tools/ubench > gpurun_out/ubench_r1b.txt 2>&1
export DY4_FRONTEND=$V WHOLE=1 S=256 NB=${NB:-47} REP=2
python tools/profile_run.py || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_frontend -s 1 -c 1 -f -o gpurun_out/prof_fe_${V} python tools/profile_run.py > gpurun_out/ncu_fe_${V}.log 2>&1
tail -3 gpurun_out/ncu_fe_${V}.log
