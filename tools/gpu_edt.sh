#!/bin/bash
# EDT correctness subset + timing, per-kernel split by ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "edt or field or signed or df_stream or shape" 2>&1 | tail -3
python tools/edt_rate.py C4; python tools/edt_rate.py C2
for v in "$@"; do echo "--- $v"; env $v python tools/edt_rate.py C4; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/edt_launches.csv python tools/edt_rate.py C4 4 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/edt_launches.csv")) if len(r) > 10]
hdr = rows[0]; k = hdr.index("Kernel Name"); v = hdr.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[1:]:
    try: d[r[k][:60]].append(float(r[v].replace(",", "")))
    except Exception: pass
for name, ts in d.items():
    print("%-62s n=%3d last=%8.1f us" % (name, len(ts), ts[-1] / 1000 if ts[-1] > 1000 else ts[-1]))
PY
