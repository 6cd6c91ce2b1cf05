mkdir -p gpurun_out
export SYGB200_VARIANT=${VARIANT:-2}
CMD="python bench.py --hours ${HOURS:-0.5} --steps 1 --warmup 1 --no-e2e --no-cpu ${EXTRA_ARGS}"
$CMD > gpurun_out/ncu_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-frame_warp} -s ${SKIP:-4} -c ${COUNT:-1} -f -o gpurun_out/${NAME:-prof} $CMD > gpurun_out/ncu_run.log 2>&1
tail -3 gpurun_out/ncu_plain.log; tail -5 gpurun_out/ncu_run.log; ls -la gpurun_out
