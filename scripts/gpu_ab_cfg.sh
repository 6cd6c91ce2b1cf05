# A/B of library builds on the secondary configs: in-tree libsygb200.so vs build/alt/$ALT (space separated); WHICH = bench_configs args
ALT=${ALT:?names of the libraries under build/alt}
cp sygnals_b200/libsygb200.so /tmp/lib_main.so
show() { python scripts/bench_configs.py ${WHICH:-cfg2} 2>/dev/null | python -c "
import sys, json
print('$1', ' '.join('%s%s=%.3f' % (d['config'][:4], d.get('n_fft',''), d['ms']) for d in map(json.loads, sys.stdin)))"; }
show main
for a in $ALT; do cp build/alt/$a sygnals_b200/libsygb200.so; show $a; done
cp /tmp/lib_main.so sygnals_b200/libsygb200.so; show main
