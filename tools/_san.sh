rm -f gpurun_out/sanitizer_r1.log
for tool in memcheck synccheck initcheck; do
  echo "==== compute-sanitizer --tool $tool python tools/sanitize.py" >> gpurun_out/sanitizer_r1.log
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize.py 2>&1 | grep -v "^$" | tail -n 14 >> gpurun_out/sanitizer_r1.log
done
for path in resident stream tile; do
  echo "==== compute-sanitizer --tool racecheck python tools/sanitize.py $path" >> gpurun_out/sanitizer_r1.log
  timeout 600 compute-sanitizer --tool racecheck --print-limit 12 python tools/sanitize.py $path 2>&1 | grep -v "^$" | tail -n 60 >> gpurun_out/sanitizer_r1.log
done
grep -E "====|ERROR SUMMARY|RACECHECK SUMMARY|DONE" gpurun_out/sanitizer_r1.log | head -40
