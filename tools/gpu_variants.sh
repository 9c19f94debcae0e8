#!/bin/bash
# A/B of library variants built next to the default one: gpu_variants.sh <workload> <lib>...
cd "$(dirname "$0")/.."
w=$1; shift
python tools/rate_c2.py $w
for lib in "$@"; do B200SV_LIB=$PWD/$lib python tools/rate_c2.py $w; done
