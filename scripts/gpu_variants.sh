mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in ${VARIANTS:-0 1 2 3 4}; do
  echo "== variant $v"
  SYGB200_VARIANT=$v python bench.py --hours ${HOURS:-2} --steps 3 --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    ln=ln.strip()
    if ln.startswith('{'):
        d=json.loads(ln); r=d['roofline']
        print('ms/step %.2f  frame %.2f  finalize %.2f  value %.0f  frac %.4f launches %d clocks %s' % (d['ms_per_step'], r['kernel_ms_per_step'], r['finalize_ms_per_step'], d['value'], r['frac'], d['gpu_launches'], d['clocks']))
    else: print(ln)
"
done
