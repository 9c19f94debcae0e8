#!/usr/bin/env python3
"""Instruction / stall-sample share per code region of check_kernels.cu from an .ncu-rep (source page)."""
import csv, io, subprocess, sys
rep, filt, nstates = sys.argv[1], sys.argv[2], float(sys.argv[3])
src = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout)))
kernel, lines = None, {}
for r in src:
    if not r: continue
    if r[0] == "Function Name": kernel = r[1]; continue
    if r[0] in ("File Path", "Line No") or r[0] == "": continue
    if filt not in (kernel or ""): continue
    try: line, inst, samp = int(r[0]), int(r[7]), int(r[6])
    except Exception: continue
    a = lines.setdefault(line, [0, 0]); a[0] += inst; a[1] += samp
text = open("/root/repo/moveit2_b200/csrc/check_kernels.cu").read().split("\n")
def find(pat):
    for i, l in enumerate(text):
        if pat in l: return i + 1
    return 10**9
marks = [("helpers (load/store/mul)", 1), ("fkState", find("void fkState")), ("posePoint", find("void posePoint")),
         ("loadModel", find("Smem<R> loadModel")), ("probeSphere", find("Probe probeSphere")), ("riskySphere", find("void riskySphere")),
         ("checkState: gathers", find("void checkState")), ("checkState: link-pair cull", find("if (!hit && do_intra")),
         ("checkState: level-2 + sphere pairs", find("while (mask)")), ("checkState: tail", find("hit = __any_sync(kFull, phit)")),
         ("kernel main loop", find("checkKernel(CheckArgs<FT> a)") - 1), ("other kernels", find("fkKernel(FkArgs a)") - 1)]
tot = sum(v[0] for v in lines.values()) or 1; ts = sum(v[1] for v in lines.values()) or 1
print("%d warp instructions = %.0f per state; %d stall samples" % (tot, tot / nstates, ts))
for i, (n, st) in enumerate(marks):
    en = marks[i + 1][1] if i + 1 < len(marks) else 10**9
    ii = sum(v[0] for l, v in lines.items() if st <= l < en); ss = sum(v[1] for l, v in lines.items() if st <= l < en)
    print("  %-36s lines %4d-%-5d %5.1f%% inst  %5.1f%% samples  %6.0f warp-inst/state" % (n, st, min(en, 9999), 100 * ii / tot, 100 * ss / ts, ii / nstates))
