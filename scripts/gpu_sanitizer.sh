#!/bin/bash
# compute-sanitizer runs on a B200 (under gpurun): memcheck on smoke() and on the all-kernel-families case, synccheck and
# racecheck on smoke() and on the tensor-core factorisation case.  Summaries -> gpurun_out/sanitizer_*.log
mkdir -p gpurun_out
CS="compute-sanitizer --print-limit 20"
timeout 900 $CS --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_smoke.log 2>&1
timeout 1500 $CS --tool memcheck python scripts/sanitizer_case.py > gpurun_out/sanitizer_memcheck_case.log 2>&1
timeout 900 $CS --tool synccheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_synccheck_smoke.log 2>&1
timeout 1500 $CS --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck_smoke.log 2>&1
for f in gpurun_out/sanitizer_*.log; do echo "== $f"; grep -E "ok|ERROR SUMMARY|RACECHECK SUMMARY|Error|error" $f | tail -4; done
