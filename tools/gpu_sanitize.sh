#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/sanitize_probe.py 2>&1 | tail -2
for tool in memcheck racecheck synccheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_probe.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize probe ok|Error|hazard" gpurun_out/sanitize_$tool.log | head -8
done
