#!/bin/bash
# round-2 GPU pass A: tests, bench, launch list, ncu of the gather microbenchmark
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/a_gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 600 python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/a_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/a_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-parity --no-c3 > gpurun_out/a_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gatherKernel -c 2 -o gpurun_out/a_gather \
  python tools/gather_probe.py > gpurun_out/a_gather_under_ncu.log 2>&1; echo "ncu gather rc=$?"
python tools/latency.py > gpurun_out/a_latency.json 2> gpurun_out/a_latency.err; echo "latency rc=$?"; cat gpurun_out/a_latency.json
python tools/gather_probe.py > gpurun_out/a_gather.json 2>/dev/null; cat gpurun_out/a_gather.json
