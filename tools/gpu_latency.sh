#!/bin/bash
# single-state latency from C, resident-kernel path on (default) and off, under a hard timeout
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
gcc -O2 -std=gnu99 -Iinclude -o /tmp/c_latency tools/c_latency.c -Lmoveit2_b200 -lb200sv -Wl,-rpath,$PWD/moveit2_b200 || exit 1
timeout 120 /tmp/c_latency moveit2_b200/resources/panda_spheres.urdf moveit2_b200/resources/panda.srdf > gpurun_out/p_c_latency.json; echo "rc=$?"; cat gpurun_out/p_c_latency.json
B200SV_LINGER=0 timeout 120 /tmp/c_latency moveit2_b200/resources/panda_spheres.urdf moveit2_b200/resources/panda.srdf > gpurun_out/p_c_latency_nolinger.json; echo "rc=$?"; cat gpurun_out/p_c_latency_nolinger.json
