# usage: scripts/ncu_report.sh <tag> [frames]  -- dumps raw/source CSVs of gpurun_out/<tag>.ncu-rep and prints the summaries
T=$1; F=${2:-117986}
ncu -i gpurun_out/$T.ncu-rep --page raw --csv > gpurun_out/raw_$T.csv 2>/dev/null
ncu -i gpurun_out/$T.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/src_$T.csv 2>/dev/null
python scripts/ncu_raw.py gpurun_out/raw_$T.csv
python scripts/ncu_sections.py gpurun_out/src_$T.csv $F
