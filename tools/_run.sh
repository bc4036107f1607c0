for pol in 0 1 2 3; do
  echo "== policy $pol"
  FOV360_SAT_POLICY=$pol timeout 60 ./tools/sat_trace.bin | head -1
  FOV360_SAT_POLICY=$pol timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -s 5 -c 1 ./tools/sat_trace.bin 2>&1 | grep -E "dram__|lts__|gpu__time"
done
echo "== R=16"; FOV360_SAT_BAND_ROWS=16 timeout 60 ./tools/sat_trace.bin | head -1
FOV360_SAT_BAND_ROWS=16 timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -s 5 -c 1 ./tools/sat_trace.bin 2>&1 | grep -E "dram__|lts__|gpu__time"
echo "== R=64"; FOV360_SAT_BAND_ROWS=64 timeout 60 ./tools/sat_trace.bin | head -1
FOV360_SAT_BAND_ROWS=64 timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -s 5 -c 1 ./tools/sat_trace.bin 2>&1 | grep -E "dram__|lts__|gpu__time"
