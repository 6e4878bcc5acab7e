"""profiles/<tag>_sass_excerpt.md from the in-tree library: per kernel, counts of the instructions that show the design
points (TMA bulk / tensor copies, LDGSTS, mbarrier SYNCS, reductions, shuffles, MUFU) and the first occurrence of each;
plus the instruction count of the compositing backward's per-pair loop.  usage: make_sass_excerpt.py <tag>"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
lib = os.path.join(ROOT, "gaustudio_b200", "libgsr_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
OPS = ["UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "REDG", "ATOMG", "ATOMS", "SHFL", "MUFU.EX2", "MUFU.RCP", "MUFU.SQRT", "MUFU.RSQ",
       "NANOSLEEP", "CREDUX", "VOTE", "HMMA", "UTCMMA", "UTCHMMA", "IMMA"]
FIRST = ["UBLKCP", "UTMALDG", "LDGSTS", "REDG", "SYNCS.ARRIVE", "SYNCS.PHASECHK", "NANOSLEEP"]


def short(mangled):
    d = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    d = d.replace("(anonymous namespace)::", "").replace("void ", "")
    d = re.sub(r"\((?!int\)|bool\)).*", "", d).split("::")[-1]
    return re.sub(r"\(int\)|\(bool\)", "", d).replace(" ", "")


kern, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = short(m.group(1))
        kern[cur] = []
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?;)", line)
    if m and cur:
        kern[cur].append(m.group(1).strip())

L = [f"# SASS evidence ({tag}) -- `cuobjdump -sass gaustudio_b200/libgsr_b200.so` (nvcc 12.9, sm_100a), `tools/make_sass_excerpt.py`\n",
     "What the design points look like in the shipped machine code.  No tensor-core instruction (HMMA / IMMA / UTCMMA) anywhere",
     "in the library: the path is gather / scatter + elementwise, not a contraction.\n",
     "| kernel | SASS instructions | instructions of interest (count) |", "|---|---|---|"]
tensor = 0
for k, ins in kern.items():
    c = collections.Counter()
    for i in ins:
        body = re.sub(r"^@!?U?P\d+\s+", "", i)
        for o in OPS:
            if body.startswith(o):
                c[o] += 1
    tensor += c["HMMA"] + c["UTCMMA"] + c["UTCHMMA"] + c["IMMA"]
    L.append(f"| `{k}` | {len(ins)} | " + (", ".join(f"{o} {c[o]}" for o in OPS if c[o]) or "-") + " |")
L.append(f"\ntensor-core instructions in the whole library: {tensor}\n")
L.append("First occurrence of each mechanism in the kernels that use it:\n\n```")
for k, ins in kern.items():
    for o in FIRST:
        for i in ins:
            if re.sub(r"^@!?U?P\d+\s+", "", i).startswith(o):
                L.append(f"{k:44s} {i}")
                break
L.append("```\n")
# per-pair loop of the compositing backward: from the first LDS.128 of a record to the loop's back edge
for k, ins in kern.items():
    if k.startswith("k_render_bwd<1"):
        red = max(i for i, s in enumerate(ins) if "REDG" in s)  # the last of the (two, predicated) reductions
        first = max(i for i, s in enumerate(ins[:red]) if s.startswith("UBREV"))  # __ffs of the survivor mask: top of the loop
        L.append(f"`{k}`: the per-(8x4 block, Gaussian) pair loop spans {red - first + 1} instructions from the survivor mask's "
                 f"`__ffs` to the `REDG` ({sum('SHFL' in s for s in ins[first:red])} SHFL, "
                 f"{sum(s.startswith('FSEL') for s in ins[first:red])} FSEL, "
                 f"{sum(s.startswith(('FFMA', 'FMUL', 'FADD')) for s in ins[first:red])} FFMA/FMUL/FADD, "
                 f"{sum('MUFU' in s for s in ins[first:red])} MUFU).\n")
open(os.path.join(ROOT, "profiles", f"{tag}_sass_excerpt.md"), "w").write("\n".join(L) + "\n")
print("\n".join(L[:40]))
