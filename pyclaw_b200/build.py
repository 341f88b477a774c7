"""Builds the CUDA library in-tree with nvcc for sm_100a.

    python -m pyclaw_b200.build [--force] [-v] [--fma | --all]

Two builds of the same sources:
  libclawb200.so      -fmad=false : strict IEEE, bit-for-bit agreement with the reference's
                      Fortran (compiled without FMA contraction) -- the default, and the one every
                      parity test runs;
  libclawb200_fma.so  -fmad=true  : nvcc may contract a*b+c into one DFMA (fewer FP64-pipe
                      instructions); results differ from the strict build at round-off level
                      (measured in profiles/README.md), selected with solver.arithmetic = 'fma'.
-lineinfo keeps the ncu source page usable.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libclawb200.so")
LIB_FMA = os.path.join(CSRC, "libclawb200_fma.so")
# translation units (compiled in parallel, linked into one shared library) and what each includes
_COMMON = ["arith.cuh", "rp.cuh", "classic.cuh", "launch.cuh", os.path.join("..", "..", "include", "clawb200.h")]
UNITS = {
    "clawb200.cu": _COMMON,
    "sweep_euler_x.cu": _COMMON,
    "sweep_euler_y.cu": _COMMON,
    "sweep_sphere.cu": _COMMON,
    "sweep_misc.cu": _COMMON,
    "step1.cu": _COMMON,
    "rp_point.cu": _COMMON,
    "sharpclaw.cu": _COMMON + ["sharpclaw.cuh"],
}
SOURCES = list(UNITS)
HEADERS = sorted({h for hs in UNITS.values() for h in hs})

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _stale(lib):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    for f in SOURCES + HEADERS:
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return False


def _obj_stale(obj, src):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in [src] + UNITS[src])


def build(force=False, verbose=False, fma=False, jobs=None):
    """Compile the translation units (object files under csrc/_obj/, only the stale ones unless
    `force`) with up to `jobs` nvcc processes, then link the shared library."""
    from concurrent.futures import ThreadPoolExecutor
    lib = LIB_FMA if fma else LIB
    if not force and not _stale(lib):
        return lib
    nvcc = os.environ.get("NVCC", "nvcc")
    flags = NVCC_FLAGS + (["-fmad=true", "-DCLAWB200_FMA=1"] if fma else ["-fmad=false"])
    if verbose:
        flags = flags + ["-Xptxas", "-v"]
    objdir = os.path.join(CSRC, "_obj", "fma" if fma else "strict")
    os.makedirs(objdir, exist_ok=True)
    objs, todo = [], []
    for src in SOURCES:
        obj = os.path.join(objdir, src[:-3] + ".o")
        objs.append(obj)
        if force or _obj_stale(obj, src):
            todo.append((src, obj))

    def compile_one(job):
        src, obj = job
        r = subprocess.run([nvcc] + flags + ["-c", "-o", obj, src], cwd=CSRC, capture_output=True, text=True)
        return src, r
    jobs = jobs or min(len(todo), os.cpu_count() or 1) or 1
    failed = None
    with ThreadPoolExecutor(max_workers=jobs) as ex:
        for src, r in ex.map(compile_one, todo):
            if verbose or r.returncode != 0:
                sys.stderr.write("==== %s ====\n%s%s" % (src, r.stdout, r.stderr))
            if r.returncode != 0:
                failed = src
    if failed:
        raise subprocess.CalledProcessError(1, "nvcc -c " + failed)
    # link under a temporary name and rename: a reader (a gpurun snapshot, a running process that
    # dlopens the library) sees either the old or the new file, never a half-written one
    subprocess.check_call([nvcc, "-shared", "-o", lib + ".tmp"] + objs, cwd=CSRC)
    os.replace(lib + ".tmp", lib)
    return lib


def build_all(force=False):
    """Both builds, side by side (two nvcc processes)."""
    import threading
    out = {}
    ts = [threading.Thread(target=lambda f=f: out.__setitem__(f, build(force=force, fma=f))) for f in (False, True)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    return out[False], out[True]


if __name__ == "__main__":
    if "--all" in sys.argv:
        print(*build_all(force="--force" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, fma="--fma" in sys.argv))
