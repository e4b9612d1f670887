"""Builds an alternative copy of the library with extra nvcc defines, for A/B runs on a GPU box:

    python tools/build_variant.py burst16 -DRL_BLOCK_BURST=16        # -> tools/_trace/librangelib_b200_burst16.so
    RL_B200_LIB=tools/_trace/librangelib_b200_burst16.so python tools/tune_fused.py

(`RL_B200_LIB` makes range_libc_b200.cabi load that file instead of the product library; tools/_trace/ is
git-ignored but travels with gpurun.)  Compile-time knobs: RL_BLOCK_BURST, RL_RM_BURST_PAIRS, RL_FUSED_GROUP_RAYS,
RL_COOP_PROBES / RL_COOP_SPACING, RL_QB, RL_TRACE."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from range_libc_b200 import build as b  # noqa: E402


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    tag, defines = sys.argv[1], sys.argv[2:]
    out = os.path.join(ROOT, "tools", "_trace")
    os.makedirs(out, exist_ok=True)
    objs, procs = [], []
    for s in b.SOURCES:
        obj = os.path.join(out, "%s_%s" % (tag, s.replace(".cu", ".o")))
        objs.append(obj)
        procs.append(subprocess.Popen([b.nvcc()] + b.NVCC_FLAGS + defines + ["-c", os.path.join(b.CSRC, s), "-o", obj]))
    if any(p.wait() != 0 for p in procs):
        raise SystemExit("nvcc failed")
    lib = os.path.join(out, "librangelib_b200_%s.so" % tag)
    subprocess.check_call([b.nvcc(), "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                                      "-lcudart_static", "-lpthread", "-ldl", "-lrt"])
    for o in objs:
        os.remove(o)
    print(lib)


if __name__ == "__main__":
    main()
