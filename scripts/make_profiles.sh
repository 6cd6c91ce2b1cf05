# usage: scripts/make_profiles.sh <tag> -- turns the outputs of `TAG=<tag> bash scripts/gpu_final.sh` (gpurun_out/) into the tracked
# summaries under profiles/ and refreshes profiles/traffic.json from the frame-kernel capture
T=${1:?tag}
G=gpurun_out; P=profiles
FR=311400   # frames of one launch of `bench.py --hours 0.5` (1800 segments x 173 frames)
for f in bench bench_reference bench_cfg1 bench_cfg2 bench_cfg3 bench_cfg5; do [ -s $G/${T}_$f.json ] && grep '^{' $G/${T}_$f.json > $P/${T}_$f.json; done
[ -s $G/${T}_launches.csv ] && cp $G/${T}_launches.csv $P/${T}_launches.csv
[ -s $G/${T}_tests.log ] && cp $G/${T}_tests.log $P/${T}_gpu_tests_tail.txt
[ -s $G/${T}_smoke.log ] && cp $G/${T}_smoke.log $P/${T}_smoke.log
if [ -s $G/${T}_warp.ncu-rep ]; then
  ncu -i $G/${T}_warp.ncu-rep --page raw --csv > $G/raw_${T}_warp.csv 2>/dev/null
  ncu -i $G/${T}_warp.ncu-rep --page source --print-source cuda,sass --csv > $G/src_${T}_warp.csv 2>/dev/null
  { echo "# $T: frame kernel (cfg4), one launch = 1800 segments = $FR frames"
    echo "# command: TAG=$T bash scripts/gpu_final.sh (ncu --set full --clock-control none --import-source on -k regex:frame_warp -s 2 -c 1; bench.py --hours 0.5 --steps 1 --warmup 1 --no-e2e --no-cpu --no-weak)"
    python scripts/ncu_raw.py $G/raw_${T}_warp.csv
    echo; echo "# executed warp-instructions per frame by opcode"
    python scripts/ncu_ops.py $G/src_${T}_warp.csv $FR
    echo; echo "# shared-memory wavefronts by source line"
    python scripts/ncu_smem_lines.py $G/src_${T}_warp.csv | head -24
    echo; echo "# hottest source lines"
    python scripts/ncu_lines.py $G/src_${T}_warp.csv 24
  } > $P/${T}_warp_ncu_summary.txt
  python - "$G/raw_${T}_warp.csv" "$T" <<'PY'
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
names = rows[0]; units = rows[1]; vals = rows[2]
def get(n):
    i = names.index(n); v = float(vals[i].replace(",", "")); u = units[i]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
d = {"workload": "cfg4", "kernel": vals[names.index("Kernel Name")], "units_in_launch": 1800, "frames_in_launch": 311400,
     "dram_bytes_read": int(get("dram__bytes_read.sum")), "dram_bytes_write": int(get("dram__bytes_write.sum")),
     "algorithmic_bytes": 1800 * (44100 * 4 + 24 * 173 * 4),
     "source": "profiles/%s_warp_ncu_summary.txt (ncu --set full --clock-control none, one launch of bench.py --hours 0.5)" % sys.argv[2]}
json.dump(d, open("profiles/traffic.json", "w"), indent=1)
print(d)
PY
fi
{ echo "# $T: STFT kernels (scripts/stft_once.py <n_fft>: 4096 x 1 s clips, magnitude; ncu --set full --clock-control none, one launch each)"
  for n in 512 2048 4096 8192; do
    [ -s $G/${T}_stft$n.ncu-rep ] || continue
    ncu -i $G/${T}_stft$n.ncu-rep --page raw --csv > $G/raw_${T}_stft$n.csv 2>/dev/null
    ncu -i $G/${T}_stft$n.ncu-rep --page source --print-source cuda,sass --csv > $G/src_${T}_stft$n.csv 2>/dev/null
    echo; echo "## n_fft $n"; python scripts/ncu_raw.py $G/raw_${T}_stft$n.csv
    echo "# shared-memory wavefronts by source line"; python scripts/ncu_smem_lines.py $G/src_${T}_stft$n.csv | head -10
  done
} > $P/${T}_stft_ncu_summary.txt
ls -la $P | grep $T
