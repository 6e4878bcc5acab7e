"""Builds experiment variants of the library (extra -D flags) into gaustudio_b200/variants/libgsr_<name>.so; a variant is
selected at run time with GSR_LIB=<path> (see tools/coresidency_ab.sh).  Test / measurement infrastructure only."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaustudio_b200 import build as B  # noqa: E402

VARIANTS = {
    "pre64": ["-DGSR_PRE_THREADS=64"],
    "pre256": ["-DGSR_PRE_THREADS=256"],
    "scatter128": ["-DGSR_SCATTER_THREADS=128"],
    "scatter512": ["-DGSR_SCATTER_THREADS=512"],
    "exactexp": ["-DGSR_BWD_FASTEXP=0"],
    "bwd48": ["-DGSR_BWD_BOUND_EXTRA=32"],
    "tiny64": ["-DGSR_SORT_THREADS_TINY=64"],
    "tiny256": ["-DGSR_SORT_THREADS_TINY=256"],
    "tinycap1024": ["-DGSR_SORT_CAP_TINY=1024"],
    "tinycap4096": ["-DGSR_SORT_CAP_TINY=4096"],
    "tinycap4096_256": ["-DGSR_SORT_CAP_TINY=4096", "-DGSR_SORT_THREADS_TINY=256"],
    "tinycap1536": ["-DGSR_SORT_CAP_TINY=1536"],
    # experimental 8-CTA cluster tile scan compiled in (then GSR_SCAN_CLUSTER=1 selects it at run time)
    "cluster_scan": ["-DGSR_WITH_CLUSTER_SCAN=1"],
    "cluster_scan1024": ["-DGSR_WITH_CLUSTER_SCAN=1", "-DGSR_SCAN_CL_THREADS=1024"],
}


def build(names=None):
    out_dir = os.path.join(B.HERE, "variants")
    os.makedirs(out_dir, exist_ok=True)
    for name, defs in VARIANTS.items():
        if names and name not in names:
            continue
        lib = os.path.join(out_dir, f"libgsr_{name}.so")
        objs, procs = [], []
        for s in B.SOURCES:
            o = os.path.join(out_dir, s.replace(".cu", f".{name}.o"))
            objs.append(o)
            procs.append((s, subprocess.Popen([B.NVCC, *B.FLAGS, *defs, "-c", os.path.join(B.CSRC, s), "-o", o],
                                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        log = []
        for s, p in procs:
            out, _ = p.communicate()
            log.append(out)
            if p.returncode != 0:
                sys.stderr.write(out)
                raise RuntimeError(f"nvcc failed on {s} ({name})")
        subprocess.check_call([B.NVCC, "-shared", "-Xlinker", f"-soname=libgsr_{name}.so", "-o", lib, *objs, "-lcudart", "-ldl"])
        for o in objs:
            os.remove(o)
        open(os.path.join(out_dir, f"ptxas_{name}.log"), "w").write("\n".join(log))
        print("built", lib)


if __name__ == "__main__":
    build(sys.argv[1:])
