#!/bin/bash
# The long parity campaign with the round's final kernels: validity vs the oracle on identical fields.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python tools/parity_10m.py 100000000 C2 > gpurun_out/r2_parity_100m_states_c2.json 2> gpurun_out/parity_c2.err; echo "C2 rc=$?"; tail -c 600 gpurun_out/r2_parity_100m_states_c2.json
timeout 900 python tools/parity_10m.py 50000000 C3 > gpurun_out/r2_parity_50m_states_c3.json 2> gpurun_out/parity_c3.err; echo "C3 rc=$?"; tail -c 400 gpurun_out/r2_parity_50m_states_c3.json
timeout 900 python tools/parity_10m.py 20000000 C3D > gpurun_out/r2_parity_20m_states_c3d.json 2> gpurun_out/parity_c3d.err; echo "C3D rc=$?"; tail -c 400 gpurun_out/r2_parity_20m_states_c3d.json
timeout 900 python tools/parity_10m.py 20000000 C3W > gpurun_out/r2_parity_20m_states_c3w.json 2> gpurun_out/parity_c3w.err; echo "C3W rc=$?"; tail -c 400 gpurun_out/r2_parity_20m_states_c3w.json
