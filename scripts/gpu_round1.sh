mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv
nproc
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r1_tests.log
cat gpurun_out/r1_tests.log
python __graft_entry__.py --smoke > gpurun_out/r1_smoke.log 2>&1; tail -5 gpurun_out/r1_smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; cat gpurun_out/r1_bench.json; tail -5 gpurun_out/r1_bench.err
python bench.py --steps 3 --warmup 3 --ws-mb 4096 --no-e2e --no-cpu > gpurun_out/r1_bench_ws4g.json 2>&1; cat gpurun_out/r1_bench_ws4g.json
python bench.py --steps 1 --warmup 1 --hours 0.5 --no-e2e --no-cpu > gpurun_out/r1_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'frame_kernel|finalize_kernel|welch' -c 200 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 1 --warmup 1 --hours 0.5 --no-e2e --no-cpu > gpurun_out/r1_ncu.log 2>&1
tail -3 gpurun_out/r1_ncu.log; tail -5 gpurun_out/r1_launches.csv
