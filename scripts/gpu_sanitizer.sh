mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitizer_case.py > gpurun_out/sanitizer_case.log 2>&1
echo "exit $?"; grep -E "ERROR SUMMARY|Invalid|sanitizer case|rror" gpurun_out/sanitizer_case.log | head -20
