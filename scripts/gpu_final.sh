# end-of-round evidence run on one B200: GPU parity tests, smoke, the default bench line, the reference arm, the secondary workloads,
# the ncu launch list of a short bench command and ncu --set full captures of the frame kernel and of the STFT kernels.
# TAG names the outputs under gpurun_out/ (copy the summaries to profiles/).
TAG=${TAG:-r02_z}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/${TAG}_tests.log; tail -1 gpurun_out/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cut -c1-200 gpurun_out/${TAG}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; cut -c1-200 gpurun_out/${TAG}_bench_reference.json
for wl in cfg1 cfg2 cfg3 cfg5; do python bench.py --workload $wl --no-cpu > gpurun_out/${TAG}_bench_${wl}.json 2>> gpurun_out/${TAG}_bench.err; cut -c1-160 gpurun_out/${TAG}_bench_${wl}.json; done
CMD="python bench.py --steps 2 --warmup 3 --hours 1 --no-e2e --no-cpu --no-weak"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"frame_warp|finalize|aggregate|pcm|time_extra" -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu.log 2>&1
CMD="python bench.py --hours 0.5 --steps 1 --warmup 1 --no-e2e --no-cpu --no-weak"
$CMD > gpurun_out/ncu_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:frame_warp -s 2 -c 1 -f -o gpurun_out/${TAG}_warp $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_full.log
for n in 512 2048 4096 8192; do
  CMD="python scripts/stft_once.py $n"
  $CMD > gpurun_out/ncu_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"stft_ring|stft_big" -s 2 -c 1 -f -o gpurun_out/${TAG}_stft$n $CMD > gpurun_out/ncu_run.log 2>&1
  tail -1 gpurun_out/ncu_run.log
done
