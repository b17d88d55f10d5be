#!/usr/bin/env python
"""Attribute the per-instruction counters of an ncu capture to source lines of the pool kernel.

The scene-specialised kernel is compiled in memory by NVRTC, so ncu cannot import its source.  NVRTC is
deterministic, though: dump the same cubin here (RTPBR_JIT_DUMP=<prefix>, no GPU needed), disassemble it with
line info and join on the instruction index.

    RTPBR_JIT_DUMP=/tmp/k python -c "..."            # writes /tmp/k.cu, /tmp/k.cubin   (see --dump)
    ncu -i rep.ncu-rep --page source --csv --print-source sass > sass.csv
    python tools/ncu_lines.py sass.csv /tmp/k.cubin [--by outer|inner] [--top 40]
"""
import argparse
import collections
import csv
import re
import subprocess

ap = argparse.ArgumentParser()
ap.add_argument("sass_csv")
ap.add_argument("cubin")
ap.add_argument("--by", default="outer", choices=["outer", "inner"])
ap.add_argument("--top", type=int, default=50)
ap.add_argument("--frame", default="pool_kernel.cuh", help="outer attribution: last frame in this file")
ap.add_argument("--within", default=None, help="file:line -- only instructions inlined under this frame, keyed by the frame just inside it")
a = ap.parse_args()

rows = list(csv.reader(open(a.sass_csv)))
hdr, data = rows[1], rows[2:]
iex, ith, ism, isrc = (hdr.index(k) for k in ("Instructions Executed", "Thread Instructions Executed", "# Samples", "Source"))

txt = subprocess.run(["nvdisasm", "--print-line-info-inline", a.cubin], capture_output=True, text=True).stdout
insts = []          # (opcode text, [frames innermost..outermost])
frames = []
pending_reset = True
in_kernel = False
for line in txt.splitlines():
    if line.startswith("//---") and ".text." in line:
        in_kernel = "k_pathtrace" in line
        continue
    if not in_kernel:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        if pending_reset:
            frames = []
            pending_reset = False
        frames.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", line)
    if m:
        insts.append((m.group(2).strip(), list(frames)))
        pending_reset = True
if len(insts) != len(data):
    raise SystemExit(f"instruction count mismatch: ncu {len(data)} vs cubin {len(insts)}")

agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for (op, fr), r in zip(insts, data):
    e, t, s = int(r[iex]), int(r[ith]), int(r[ism])
    if a.within:
        wf, wl = a.within.split(":")
        pos = [i for i, f in enumerate(fr) if f[0] == wf and f[1] == int(wl)]
        if not pos:
            continue
        i = pos[-1]            # frames are innermost first
        key = fr[i - 1] if i > 0 else fr[i]
    elif a.by == "inner":
        key = fr[0] if fr else ("?", 0)
    else:
        cand = [f for f in fr if f[0] == a.frame]
        key = cand[-1] if cand else (fr[-1] if fr else ("?", 0))
    for k, v in zip(range(3), (e, t, s)):
        agg[key][k] += v
        tot[k] += v
print(f"total: {tot[0] / 1e9:.3f} G warp-inst, {tot[1] / max(tot[0], 1):.2f} threads/inst, {tot[2]} samples")
print(f"{'where':32s} {'inst%':>7s} {'lanes':>6s} {'samples%':>8s}")
for key, (e, t, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[: a.top]:
    print(f"{key[0] + ':' + str(key[1]):32s} {100 * e / tot[0]:7.2f} {t / max(e, 1):6.1f} {100 * s / max(tot[2], 1):8.2f}")
