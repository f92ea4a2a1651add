# compute-sanitizer passes over the small cases (one B200); output -> gpurun_out/r02_sanitizer.txt
mkdir -p gpurun_out; out=gpurun_out/r02_sanitizer.txt; : > $out
for which in short long longer; do
  for tool in memcheck racecheck synccheck; do
    echo "== $which $tool" >> $out
    timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_small.py $which 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|status|hazard|Error|error" | head -12 >> $out
  done
done
cat $out
