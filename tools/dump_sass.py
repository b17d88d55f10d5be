#!/usr/bin/env python
"""SASS evidence for profiles/ (no GPU needed: NVRTC and nvdisasm run here, and NVRTC is deterministic).

Compiles the scene-specialised translation unit of a preset with the build options rtpbr_pathtrace uses
(rtpbr_jit_compile_check + RTPBR_JIT_DUMP) and writes
  * the march loop of k_pathtrace_pool_jit -- the innermost loop that holds the VOTE and the MUFU.RSQ / packed FFMA2
    instructions of the scene evaluation -- with an opcode histogram, and
  * (bunny scenes) the out-of-line MLP routines.

    python tools/dump_sass.py cornell_box_shortest 1024 1024 profiles/r02_sass_c1_march_loop.txt
    python tools/dump_sass.py bunny_glass 1024 1024 profiles/r02_sass_c2_bunny_mlp.txt --function sd_bunny_mlp --function sin4_rt
"""
import argparse
import collections
import os
import re
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raytracingpbr_b200 import _native, scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("preset")
ap.add_argument("width", type=int)
ap.add_argument("height", type=int)
ap.add_argument("out")
ap.add_argument("--function", action="append", default=[], help="also list this out-of-line device function")
a = ap.parse_args()

cfg, objs, cam, tm = getattr(scenes, a.preset)(a.width, a.height)
prefix = os.path.join(tempfile.mkdtemp(), "k")
os.environ["RTPBR_JIT_DUMP"] = prefix
_native.jit_compile_check(cfg, [o.to_native() for o in objs])
usage = subprocess.run(["cuobjdump", "-res-usage", prefix + ".cubin"], capture_output=True, text=True).stdout
sass = subprocess.run(["nvdisasm", "-c", prefix + ".cubin"], capture_output=True, text=True).stdout.split("\n")
src = open(prefix + ".cu").read()

INST = re.compile(r"/\*[0-9a-f]{4,}\*/\s+(.*?);")


def section(name_part):
    """lines of the .text section whose name contains name_part"""
    out, on = [], False
    for line in sass:
        if line.startswith("//---") and ".text." in line:
            on = name_part in line
        elif line.startswith("//---"):
            on = False
        if on:
            out.append(line)
    return out


def histogram(lines):
    h = collections.Counter()
    for line in lines:
        m = INST.search(line)
        if m:
            op = m.group(1).split()[0]
            if op.startswith("@"):
                op = m.group(1).split()[1]
            h[op.split(".")[0]] += 1
    return h


kernel = section("k_pathtrace_pool_jit")
labels = {}
for i, line in enumerate(kernel):
    m = re.match(r"^(\.L_x_\d+):", line.strip())
    if m:
        labels[m.group(1)] = i
loops = []
for i, line in enumerate(kernel):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?`\((\.L_x_\d+)\)", line)
    if m and m.group(1) in labels and labels[m.group(1)] < i:
        body = kernel[labels[m.group(1)]:i + 1]
        txt = "\n".join(body)
        if "VOTE" in txt and "MUFU" in txt:
            loops.append((len([b for b in body if INST.search(b)]), labels[m.group(1)], i))
loops.sort()
with open(a.out, "w") as f:
    f.write(f"# {a.preset} {a.width}x{a.height}: scene-specialised kernel as rtpbr_pathtrace builds it (tools/dump_sass.py; NVRTC, sm_100a)\n")
    f.write("# " + " ".join(l.strip() for l in usage.strip().split("\n")[-2:]) + "\n")
    f.write("# build switches: " + " ".join(l for l in src.split("\n") if l.startswith("#define RT_JIT") or l.startswith("#define RT_RESOLVE")) + "\n")
    if loops:
        n, lo, hi = loops[0]
        body = kernel[lo:hi + 1]
        h = histogram(body)
        f.write(f"# MARCH LOOP: {n} instructions per sphere-tracing step (one scene evaluation + vote); "
                f"local-memory traffic: {h.get('LDL', 0)} LDL / {h.get('STL', 0)} STL\n")
        f.write("# opcodes: " + ", ".join(f"{k} {v}" for k, v in h.most_common()) + "\n")
        f.write("\n".join(re.sub(r"\s*/\*[0-9a-f]{16}\*/\s*$", "", b).rstrip() for b in body) + "\n")
    for fn in a.function:
        # out-of-line device functions live in the kernel's section: from their `$...<name>...:` label to the next `.type`
        body, on = [], False
        for line in kernel:
            if re.match(r"\$\S*" + re.escape(fn) + r"\S*:\s*$", line.strip()):
                on = True
            elif on and re.match(r"\s*\.type\s", line):
                break
            elif on:
                body.append(line)
        h = histogram(body)
        f.write(f"\n# FUNCTION {fn}: {sum(h.values())} instructions; opcodes: " + ", ".join(f"{k} {v}" for k, v in h.most_common()) + "\n")
        f.write("\n".join(re.sub(r"\s*/\*[0-9a-f]{16}\*/\s*$", "", b).rstrip() for b in body if INST.search(b) or b.strip().startswith(".L_")) + "\n")
print(open(a.out).read()[:1200])
