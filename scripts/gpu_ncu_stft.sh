# ncu --set full captures of the STFT kernels: NFFTS="256 512 2048" TAG=r02_d
mkdir -p gpurun_out
for n in ${NFFTS:-512}; do
  CMD="python scripts/stft_once.py $n"
  $CMD > gpurun_out/ncu_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"${KERNEL:-stft_ring|frame_warp|stft_tile}" -s 2 -c 1 -f -o gpurun_out/${TAG:-r02}_stft$n $CMD > gpurun_out/ncu_run.log 2>&1
  tail -1 gpurun_out/ncu_run.log
done
