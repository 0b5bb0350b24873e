#!/bin/bash
# tcgen05 factorisation pass: kernel self-tests + factorisation parity, then factorisation wall times with the tensor-core
# GEMMs off / on (profile_setup.py prints ms per segp_set_model + segp_factorize), then launch lists.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fact_i8.py -m gpu -q -x -s > gpurun_out/pytest_fact.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fact.log; grep -E "^n=|passed|failed|rc=" gpurun_out/pytest_fact.log
for cfg in C4 C5; do for mode in 0 1; do for slots in 4 1; do
  SEGP_FACT_SLOTS=$slots SEGP_FACT_I8=$mode timeout 600 python scripts/profile_setup.py $cfg 4 > gpurun_out/setup_${cfg}_f${mode}_s$slots.log 2>&1; echo "slots $slots: $(tail -2 gpurun_out/setup_${cfg}_f${mode}_s$slots.log | tr '\n' ' ')"
done; done; done
SEGP_FACT_I8=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_setup_c4_f1.csv python scripts/profile_setup.py C4 0 > gpurun_out/ncu_setup_f1.log 2>&1
SEGP_FACT_I8=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_setup_c5_f1.csv python scripts/profile_setup.py C5 0 > gpurun_out/ncu_setup_c5_f1.log 2>&1
