"""Builds librangelib_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m range_libc_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librangelib_b200.so")
SOURCES = ["rl_abi.cu", "rl_occ.cu", "rl_edt.cu", "rl_cddt.cu", "rl_cast.cu", "rl_sort.cu", "rl_pf.cu"]
HEADERS = ["rl_internal.cuh", "rl_math.cuh", os.path.join(ROOT, "include", "rangelib_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # parity: no FP contraction anywhere (see DESIGN.md "Numerics"); IEEE div/sqrt
    "--fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-O2",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(HERE, "build", s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs + [os.path.abspath(__file__)]):
            cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("---- %s ----\n%s\n" % (s, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static",
                                                         "-lpthread", "-ldl", "-lrt"]
        subprocess.check_call(cmd)
    return LIB


def build_cython(force=False):
    """Builds the drop-in ``range_libc`` extension module (pywrapper/RangeLibc.pyx) against the C ABI."""
    import sysconfig
    pyx = os.path.join(HERE, "pywrapper", "RangeLibc.pyx")
    csrc = os.path.join(HERE, "build", "RangeLibc.c")
    ext = os.path.join(HERE, "pywrapper", "range_libc" + sysconfig.get_config_var("EXT_SUFFIX"))
    if not (force or _stale(ext, [pyx, LIB, os.path.join(ROOT, "include", "rangelib_b200.h")])):
        return ext
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    subprocess.check_call([sys.executable, "-m", "cython", "-3", "--module-name", "range_libc", pyx, "-o", csrc])
    inc = sysconfig.get_paths()["include"]
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-w", "-I", inc, "-I", os.path.join(ROOT, "include"), csrc,
                           "-o", ext, "-L", HERE, "-lrangelib_b200", "-Wl,-rpath,$ORIGIN/.."])
    return ext


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_cython(force="--force" in sys.argv))
