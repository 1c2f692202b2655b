#!/bin/bash
# compute-sanitizer over scripts/sanitizer_probe.py (every production kernel at a small ragged batch)
O=gpurun_out/r02san; mkdir -p $O
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitizer_probe.py > $O/$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitizer probe ok' $O/$tool.log | tr '\n' ' ')"
done
