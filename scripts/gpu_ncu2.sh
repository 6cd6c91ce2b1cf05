mkdir -p gpurun_out
export SYGB200_TWO_STAGE=1
CMD="python bench.py --hours 0.5 --steps 1 --warmup 1 --no-e2e --no-cpu --ws-mb 2048"
$CMD > gpurun_out/ncu_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:frame_warp -s 2 -c 2 -f -o gpurun_out/$NAME $CMD > gpurun_out/ncu_run.log 2>&1
tail -2 gpurun_out/ncu_run.log
