#!/usr/bin/env python3
"""Instruction / stall-sample share per function region of check_kernels.cu from an .ncu-rep (source page).
Usage: tools/ncu_regions.py rep kernel-substring n_states [top_lines]"""
import csv, io, subprocess, sys
rep, filt, nstates = sys.argv[1], sys.argv[2], float(sys.argv[3])
ntop = int(sys.argv[4]) if len(sys.argv) > 4 else 20
src = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout)))
hdr = next(r for r in src if r and r[0] == "Line No")
iThr, iInst, iS = hdr.index("Thread Instructions Executed"), hdr.index("Instructions Executed"), hdr.index("# Samples")
kernel, lines = None, {}
for r in src:
    if not r: continue
    if r[0] == "Function Name": kernel = r[1]; continue
    if r[0] in ("File Path", "Line No") or r[0] == "": continue
    if filt not in (kernel or ""): continue
    try: line, inst, samp, thr = int(r[0]), int(r[iInst]), int(r[iS]), int(r[iThr])
    except Exception: continue
    a = lines.setdefault(line, [0, 0, 0]); a[0] += inst; a[1] += samp; a[2] += thr
text = open("/root/repo/moveit2_b200/csrc/check_kernels.cu").read().split("\n")
def find(pat):
    for i, l in enumerate(text):
        if pat in l: return i + 1
    return 10**9
marks = sorted([("helpers (load/store/mul)", 1), ("fkState", find("void fkState")), ("posePoint", find("void posePoint")),
         ("loadModel", find("Smem<R> loadModel")), ("riskySphere", find("void riskySphere")), ("sphereChunk", find("void sphereChunk")),
         ("pairCull", find("void pairCull")), ("pairItem", find("void pairItem")), ("kernel main loop", find("constexpr int kQueueCap")),
         ("other kernels", find("fkKernel(FkArgs a)") - 1)], key=lambda m: m[1])
tot = sum(v[0] for v in lines.values()) or 1; ts = sum(v[1] for v in lines.values()) or 1
print("%d warp instructions = %.0f per 32-state tile (%.0f per state); %d stall samples" % (tot, 32 * tot / nstates, tot / nstates, ts))
for i, (n, st) in enumerate(marks):
    en = marks[i + 1][1] if i + 1 < len(marks) else 10**9
    sel = [v for l, v in lines.items() if st <= l < en]
    ii, ss, tt = sum(v[0] for v in sel), sum(v[1] for v in sel), sum(v[2] for v in sel)
    print("  %-26s %5.1f%% inst  %5.1f%% samples  %7.0f warp-inst/tile  avg active lanes %4.1f" % (n, 100 * ii / tot, 100 * ss / ts, 32 * ii / nstates, tt / max(ii, 1)))
for l, v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:ntop]:
    print("  line %4d %5.1f%% inst %5.1f%% samp lanes %4.1f  %s" % (l, 100 * v[0] / tot, 100 * v[1] / ts, v[2] / max(v[0], 1), text[l - 1].strip()[:95]))
