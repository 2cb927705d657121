#!/bin/bash
# compute-sanitizer over the trait kernels with the episode's event counters switched on (the code added for ABI 10)
T=gpurun_out/r02san2
mkdir -p $T
CS=/usr/local/cuda/bin/compute-sanitizer
for v in metabolic_ev cooperation_ev investment_ev; do
  for tool in memcheck; do
    timeout 100 $CS --tool $tool --print-limit 20 python scripts/sanitize_rollout.py $v 30 64 > $T/sanitizer_${tool}_$v.log 2>&1
    echo "$tool $v rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $T/sanitizer_${tool}_$v.log | tail -1) | $(grep ' ok ' $T/sanitizer_${tool}_$v.log | tail -1)"
  done
done
