# one short device-timed bench line per environment variant: VARS="A=1;B=2" (";"-separated env assignments, "-" = none)
line() { env $2 python bench.py --hours ${HOURS:-2} --steps 5 --warmup 3 --no-e2e --no-cpu --no-weak --workload ${WL:-cfg4} 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); r=d['roofline']; print('$1 ms/step %.3f frame %.3f finalize %.3f value %.0f' % (d['ms_per_step'], r['kernel_ms_per_step'], r['finalize_ms_per_step'], d['value']))
"; }
for i in $(seq ${REPS:-2}); do
  IFS=';' read -ra VS <<< "${VARS:--}"
  for v in "${VS[@]}"; do if [ "$v" = "-" ]; then line default "X_=1"; else line "$v" "$v"; fi; done
done
