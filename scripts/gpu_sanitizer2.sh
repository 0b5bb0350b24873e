mkdir -p gpurun_out
for tool in synccheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 10 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "SUMMARY|smoke|hazard|Barrier error" gpurun_out/sanitizer_$tool.log | head -8
done
