#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/fp32_error.py 20000 > gpurun_out/fp32_error.json; cat gpurun_out/fp32_error.json | tr -d '\n' | cut -c1-1500; echo
python tools/rate_c2.py C2; B200SV_FP32_EPS=4e-6 python tools/rate_c2.py C2; B200SV_FP32_EPS=1e-5 python tools/rate_c2.py C2
python tools/rate_c2.py C3D; B200SV_FP32_EPS=4e-6 python tools/rate_c2.py C3D
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
