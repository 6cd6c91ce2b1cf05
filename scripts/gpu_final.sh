# end-of-session evidence run on one B200: GPU parity tests, smoke, the default bench line, the reference arm, the ncu launch list of a
# short bench command and one `ncu --set full` capture of the frame kernel.  TAG names the outputs under gpurun_out/.
TAG=${TAG:-r01_final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${TAG}_tests.log; tail -1 gpurun_out/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cut -c1-300 gpurun_out/${TAG}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; cut -c1-300 gpurun_out/${TAG}_bench_reference.json
CMD="python bench.py --steps 2 --warmup 3 --hours 1 --no-e2e --no-cpu --no-weak"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu.log 2>&1
CMD="python bench.py --hours 0.5 --steps 1 --warmup 1 --no-e2e --no-cpu --no-weak"
ncu --set full --clock-control none --import-source on -k regex:frame_warp -s 2 -c 1 -f -o gpurun_out/${TAG}_warp $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
