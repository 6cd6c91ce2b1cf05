# A/B of two builds of the library on one box: the in-tree libsygb200.so against build/alt/$ALT (an experiment build made with
# `python -m sygnals_b200.build -DNAME=VALUE --out=build/alt/<file>.so`); alternates the two REPS times to average out drift
ALT=${ALT:?names of the libraries under build/alt}
cp sygnals_b200/libsygb200.so /tmp/lib_main.so
line() { python bench.py --hours ${HOURS:-2} --steps 5 --warmup 3 --no-e2e --no-cpu --no-weak 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); r=d['roofline']; print('$1 ms/step %.3f frame %.3f finalize %.3f aggregate %.3f' % (d['ms_per_step'], r['kernel_ms_per_step'], r['finalize_ms_per_step'], r.get('aggregate_ms_per_step', 0.0)))
"; }
for i in $(seq ${REPS:-3}); do
  cp /tmp/lib_main.so sygnals_b200/libsygb200.so; line main
  for a in $ALT; do cp build/alt/$a sygnals_b200/libsygb200.so; line "$a"; done
done
cp /tmp/lib_main.so sygnals_b200/libsygb200.so
