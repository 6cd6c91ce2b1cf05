# one GPU call: parity tests, A/B of the in-tree library against build/alt/$ALT on cfg4 and on the secondary configs, then one
# ncu --set full capture of the frame kernel (NAME) -- outputs under gpurun_out/
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${TAG:-ab}_tests.log
REPS=${REPS:-2} ALT="$ALT" bash scripts/gpu_ab.sh 2>&1 | tee gpurun_out/${TAG:-ab}_cfg4.log
WHICH="${WHICH:-cfg2 cfg3 cfg5}" ALT="$ALT" bash scripts/gpu_ab_cfg.sh 2>&1 | tee gpurun_out/${TAG:-ab}_cfgs.log
if [ -n "$NAME" ]; then
  CMD="python bench.py --hours 0.5 --steps 1 --warmup 1 --no-e2e --no-cpu --no-weak"
  $CMD > gpurun_out/ncu_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-frame_warp} -s ${SKIP:-2} -c 1 -f -o gpurun_out/$NAME $CMD > gpurun_out/ncu_run.log 2>&1
  tail -2 gpurun_out/ncu_run.log
fi
