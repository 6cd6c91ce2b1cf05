# A/B of environment switches of the in-tree library on cfg4: ENVS="A=1 B=2" (each entry one VAR=value; "-" = default)
line() { env $2 python bench.py --hours ${HOURS:-2} --steps 5 --warmup 3 --no-e2e --no-cpu --no-weak 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); r=d['roofline']; print('$1 ms/step %.3f frame %.3f finalize %.3f' % (d['ms_per_step'], r['kernel_ms_per_step'], r['finalize_ms_per_step']))
"; }
for i in $(seq ${REPS:-2}); do
  line default X_=0
  for e in $ENVS; do line "$e" "$e"; done
done
