#!/bin/bash
# Regenerates the committed summaries under profiles/ from `ncu --set full` reports brought back in
# gpurun_out/ (run on the tree that was profiled: the traffic entries are stamped with source hashes).
#   tools/refresh_profiles.sh gpurun_out/prof_main.ncu-rep gpurun_out/prof_img.ncu-rep [launches.csv]
set -e
cd "$(dirname "$0")/.."
main=$1; img=$2; launches=$3; tag=${TAG:-r02}
ncu -i "$main" --page raw --csv 2>/dev/null | python tools/ncu_raw_summary.py > profiles/${tag}_main_ncu_raw_summary.txt
for k in onepass sample interpolate; do
  ncu -i "$main" --page source --csv -k regex:sat_$k 2>/dev/null > /tmp/src_$k.csv
  python tools/ncu_source_summary.py /tmp/src_$k.csv > profiles/${tag}_main_ncu_source_$k.txt 2>&1
done
ncu -i "$img" --page raw --csv 2>/dev/null | python tools/ncu_raw_summary.py > profiles/${tag}_img_ncu_raw_summary.txt
ncu -i "$img" --page source --csv -k regex:img_interpolate 2>/dev/null > /tmp/src_lp.csv
python tools/ncu_source_summary.py /tmp/src_lp.csv > profiles/${tag}_img_ncu_source_interpolate.txt 2>&1
rm -f profiles/${tag}_traffic.json
python tools/make_traffic_json.py "$main" 8k 16 profiles/${tag}_traffic.json "profiles/${tag}_main_ncu_raw_summary.txt (tools/profile_step.py --batch 16)" > /dev/null
python tools/make_traffic_json.py "$img" 8k 1 profiles/${tag}_traffic.json "profiles/${tag}_img_ncu_raw_summary.txt (tools/logpolar_stats.py --time 8k)" > /dev/null
[ -n "$launches" ] && cp "$launches" profiles/${tag}_ncu_launches_bench_8k.csv
python tools/sass_histogram.py > profiles/${tag}_sass_opcodes.txt
echo refreshed profiles/${tag}_*
