mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
SYGB200_TWO_STAGE=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
fmt='import sys, json
for ln in sys.stdin:
    ln=ln.strip()
    if ln.startswith("{"):
        d=json.loads(ln); r=d["roofline"]
        print("ms/step %.2f  frame %.2f  finalize %.2f  value %.0f  launches %d" % (d["ms_per_step"], r["kernel_ms_per_step"], r["finalize_ms_per_step"], d["value"], d["gpu_launches"]))
    else: print(ln)'
for ts in 0 1; do for ws in 64 128 256 1024; do
  echo "== two_stage=$ts ws=$ws MB"
  SYGB200_TWO_STAGE=$ts python bench.py --hours 2 --steps 3 --warmup 3 --no-e2e --no-cpu --ws-mb $ws 2>&1 | python -c "$fmt"
done; done
