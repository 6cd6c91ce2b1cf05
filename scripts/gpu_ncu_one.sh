# one ncu --set full capture: CMD="python scripts/cfg3_once.py" KERNEL=frame_warp NAME=r02_e_cfg3 SKIP=2
mkdir -p gpurun_out
$CMD > gpurun_out/ncu_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-frame_warp} -s ${SKIP:-2} -c 1 -f -o gpurun_out/$NAME $CMD > gpurun_out/ncu_run.log 2>&1
tail -1 gpurun_out/ncu_run.log
