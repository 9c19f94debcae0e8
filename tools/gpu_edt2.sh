#!/bin/bash
# round-2 EDT pass 2: correctness subset, then the tiling / PDL / scatter variants on C4 and C2, then the per-kernel split
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "edt or field or signed or df_stream or shape or compact" 2>&1 | tail -3
run() { echo "--- $*"; env "$@" python tools/edt_rate.py C4 2>&1 | tail -1; }
run B200SV_EDT_ZY=1 B200SV_EDT_PDL=0
run B200SV_EDT_PDL=0
run B200SV_EDT_PDL=1
for zs in 3 4 5 6 8 9; do run B200SV_EDT_ZS=$zs; done
run B200SV_EDT_CHUNK=200
run B200SV_EDT_CHUNK=200 B200SV_EDT_ZS=3
run B200SV_SCATTER=1
run B200SV_SCATTER=1 B200SV_SCATTER_GRID=1184
run B200SV_SCATTER_GRID=1184
run B200SV_SCATTER_GRID=4736
echo "--- C2"; B200SV_EDT_ZY=1 B200SV_EDT_PDL=0 python tools/edt_rate.py C2 | tail -1; python tools/edt_rate.py C2 | tail -1
for zs in 3 4 5; do echo "--- C2 zs=$zs"; B200SV_EDT_ZS=$zs python tools/edt_rate.py C2 | tail -1; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/edt2_launches.csv python tools/edt_rate.py C4 4 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/edt2_launches.csv")) if len(r) > 10]
hdr = rows[0]; k = hdr.index("Kernel Name"); v = hdr.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[1:]:
    try: d[r[k][:60]].append(float(r[v].replace(",", "")))
    except Exception: pass
for name, ts in d.items():
    print("%-62s n=%3d last=%8.1f us" % (name, len(ts), ts[-1] / 1000 if ts[-1] > 1000 else ts[-1]))
PY
