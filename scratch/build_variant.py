"""Build a tuning variant of the library: the Euler sweep translation units are recompiled with
extra -D flags, everything else is taken from the existing object files of the chosen build.

    python scratch/build_variant.py NAME [--fma] [--x "-DFOO=1 ..."] [--y "..."] [--units a.cu,b.cu]

Result: pyclaw_b200/csrc/libclawb200_NAME.so   (select with CLAWB200_LIB=...)."""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyclaw_b200 import build as B  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("name")
ap.add_argument("--fma", action="store_true")
ap.add_argument("--x", default="")
ap.add_argument("--y", default="")
ap.add_argument("--all", default="", help="flags for every recompiled unit")
ap.add_argument("--units", default="sweep_euler_x.cu,sweep_euler_y.cu")
ap.add_argument("-v", action="store_true")
ap.add_argument("--contract-only", action="store_true", help="-fmad=true without the relaxed division (CLAWB200_FMA)")
a = ap.parse_args()

B.build(fma=a.fma)  # make sure the base objects exist
base = os.path.join(B.CSRC, "_obj", "fma" if a.fma else "strict")
vdir = os.path.join(B.CSRC, "_obj", "var_" + a.name)
os.makedirs(vdir, exist_ok=True)
flags = B.NVCC_FLAGS + (["-fmad=true"] + ([] if a.contract_only else ["-DCLAWB200_FMA=1"]) if a.fma else ["-fmad=false"])
units = [u for u in a.units.split(",") if u]
procs = []
for u in units:
    extra = a.all.split() + (a.x.split() if u == "sweep_euler_x.cu" else a.y.split() if u == "sweep_euler_y.cu" else [])
    obj = os.path.join(vdir, u[:-3] + ".o")
    cmd = ["nvcc"] + flags + extra + (["-Xptxas", "-v"] if a.v else []) + ["-c", "-o", obj, u]
    procs.append((u, obj, subprocess.Popen(cmd, cwd=B.CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
objs = []
for u, obj, p in procs:
    out = p.communicate()[0]
    if p.returncode != 0 or a.v:
        sys.stderr.write(out)
    if p.returncode != 0:
        raise SystemExit("compile failed: " + u)
    objs.append(obj)
for src in B.SOURCES:
    if src not in units:
        objs.append(os.path.join(base, src[:-3] + ".o"))
lib = os.path.join(B.CSRC, "libclawb200_%s.so" % a.name)
subprocess.check_call(["nvcc", "-shared", "-o", lib + ".tmp"] + objs, cwd=B.CSRC)
os.replace(lib + ".tmp", lib)
print(lib)
