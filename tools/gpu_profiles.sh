#!/bin/bash
# round-2 profile pass: C latency, ncu --set full of the headline kernels (each after a plain run exited 0)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
gcc -O2 -std=gnu99 -Iinclude -o /tmp/c_latency tools/c_latency.c -Lmoveit2_b200 -lb200sv -Wl,-rpath,$PWD/moveit2_b200 && \
  /tmp/c_latency moveit2_b200/resources/panda_spheres.urdf moveit2_b200/resources/panda.srdf > gpurun_out/p_c_latency.json; cat gpurun_out/p_c_latency.json
B200SV_LATENCY=0 /tmp/c_latency moveit2_b200/resources/panda_spheres.urdf moveit2_b200/resources/panda.srdf > gpurun_out/p_c_latency_off.json; cat gpurun_out/p_c_latency_off.json
gcc -O2 -std=gnu99 -Iinclude -o /tmp/c_edge_latency tools/c_edge_latency.c -Lmoveit2_b200 -lb200sv -Wl,-rpath,$PWD/moveit2_b200 -lm && \
  /tmp/c_edge_latency moveit2_b200/resources/panda_spheres.urdf moveit2_b200/resources/panda.srdf > gpurun_out/p_c_edge_latency.json; cat gpurun_out/p_c_edge_latency.json
python tools/rate_c2.py C2 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:fastCheckKernel -c 1 -o gpurun_out/p_fast \
  python tools/rate_c2.py C2 > /dev/null 2>&1; echo "ncu fast rc=$?"
python tools/edt_rate.py C4 4 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:'scatterKernel|edtZYKernel|edtXKernelVec' -s 9 -c 3 -o gpurun_out/p_edt \
  python tools/edt_rate.py C4 4 > /dev/null 2>&1; echo "ncu edt rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:'latencyCheckKernel|fkKernel' -c 4 -o gpurun_out/p_small \
  python tools/sanitize_probe.py > /dev/null 2>&1; echo "ncu small rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/p_launches_bench.csv \
  python bench.py --steps 3 --warmup 3 --no-parity --no-c3 > gpurun_out/p_ncu_bench.log 2>&1; echo "ncu list rc=$?"
