#!/bin/bash
# one / 8 / 64 edges per b200sv_check_motions call from C: few-edges path on (default) and off
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
gcc -O2 -std=gnu99 -Iinclude -o /tmp/c_edge_latency tools/c_edge_latency.c -Lmoveit2_b200 -lb200sv -Wl,-rpath,$PWD/moveit2_b200 -lm || exit 1
timeout 120 /tmp/c_edge_latency moveit2_b200/resources/panda_spheres.urdf moveit2_b200/resources/panda.srdf > gpurun_out/p_c_edge_latency.json; echo "rc=$?"; cat gpurun_out/p_c_edge_latency.json
B200SV_EDGE_LATENCY=0 timeout 120 /tmp/c_edge_latency moveit2_b200/resources/panda_spheres.urdf moveit2_b200/resources/panda.srdf > gpurun_out/p_c_edge_latency_off.json; echo "rc=$?"; cat gpurun_out/p_c_edge_latency_off.json
