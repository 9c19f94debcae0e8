#!/usr/bin/env python3
"""Per-source-line instruction / stall-sample share of one kernel from an .ncu-rep (source page), grouped by the
source file's `// ---- region` markers and function definitions.
Usage: tools/ncu_lines.py rep kernel-substring n_states source.cu [top_lines]"""
import csv, io, re, subprocess, sys
rep, filt, nstates, srcfile = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
ntop = int(sys.argv[5]) if len(sys.argv) > 5 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
src = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in src if r and r[0] == "Line No")
iThr, iInst, iS = hdr.index("Thread Instructions Executed"), hdr.index("Instructions Executed"), hdr.index("# Samples")
kernel, fpath, lines, other = None, None, {}, [0, 0, 0]
for r in src:
    if not r: continue
    if r[0] == "Function Name": kernel = r[1]; continue
    if r[0] == "File Path": fpath = r[1]; continue
    if r[0] in ("Line No", ""): continue
    if filt not in (kernel or ""): continue
    try: line, inst, samp, thr = int(r[0]), int(r[iInst]), int(r[iS]), int(r[iThr])
    except Exception: continue
    if fpath and not fpath.endswith(srcfile.split("/")[-1]):
        other[0] += inst; other[1] += samp; other[2] += thr; continue
    a = lines.setdefault(line, [0, 0, 0]); a[0] += inst; a[1] += samp; a[2] += thr
text = open(srcfile).read().split("\n")
marks = [("(top)", 1)]
for i, l in enumerate(text):
    m = re.match(r"\s*// ---- (.*?) -*$", l) or re.match(r"^(?:template.*\n)?__device__ .*? (\w+)\(", l) or re.match(r"^__global__.*? (\w+)\(", l)
    if m: marks.append((m.group(1).strip()[:40], i + 1))
tot = sum(v[0] for v in lines.values()) + other[0] or 1; ts = sum(v[1] for v in lines.values()) + other[1] or 1
print("%d warp instructions = %.0f per 32-state tile (%.1f per state); %d stall samples" % (tot, 32 * tot / nstates, tot / nstates, ts))
for i, (n, st) in enumerate(marks):
    en = marks[i + 1][1] if i + 1 < len(marks) else 10**9
    sel = [v for l, v in lines.items() if st <= l < en]
    ii, ss, tt = sum(v[0] for v in sel), sum(v[1] for v in sel), sum(v[2] for v in sel)
    if ii: print("  %-42s %5.1f%% inst  %5.1f%% samples  %7.0f warp-inst/tile  lanes %4.1f" % (n, 100 * ii / tot, 100 * ss / ts, 32 * ii / nstates, tt / max(ii, 1)))
print("  %-42s %5.1f%% inst  %5.1f%% samples  %7.0f warp-inst/tile  lanes %4.1f" % ("(other files: libdevice, headers)", 100 * other[0] / tot, 100 * other[1] / ts, 32 * other[0] / nstates, other[2] / max(other[0], 1)))
for l, v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:ntop]:
    print("  line %4d %5.1f%% inst %5.1f%% samp lanes %4.1f  %s" % (l, 100 * v[0] / tot, 100 * v[1] / ts, v[2] / max(v[0], 1), text[l - 1].strip()[:100]))
