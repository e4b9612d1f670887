"""Per-kernel SASS opcode summary of librangelib_b200.so (cuobjdump -sass), so that the memory-path design is visible
in the repo and changes show up in review: which kernels use LDG / LDS / STS / async copies (LDGSTS, UBLKCP, UTMALDG),
cache-policy loads (LDG...EF / EL, createpolicy), warp collectives (SHFL, VOTE, REDUX / CREDUX, MATCH), barriers, the
XU conversions (F2I / I2F / MUFU) and FP64.

    python tools/sass_summary.py > profiles/r02/sass_opcodes_r02.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "range_libc_b200", "librangelib_b200.so")
GROUPS = [
    ("global loads (LDG)", r"^LDG"), ("  of which evict-first / streaming (.EF)", r"^LDG.*\.EF"),
    ("  of which 128-bit", r"^LDG.*\.128"),
    ("global stores (STG)", r"^STG"), ("shared loads (LDS)", r"^LDS"), ("shared stores (STS)", r"^STS"),
    ("async copies (LDGSTS / UBLKCP / UTMALDG / UTMASTG)", r"^(LDGSTS|UBLKCP|UTMALDG|UTMASTG)"),
    ("mbarrier / SYNCS", r"^SYNCS"), ("atomics (ATOM / ATOMS / ATOMG / RED)", r"^(ATOM|RED)"),
    ("warp shuffles (SHFL)", r"^SHFL"), ("votes (VOTE / VOTEU)", r"^VOTE"), ("warp reductions (REDUX / CREDUX)", r"^C?REDUX"),
    ("CTA barriers (BAR)", r"^BAR"),
    ("float->int (F2I)", r"^F2I"), ("int->float (I2F)", r"^I2F"), ("MUFU (rsqrt, rcp, ...)", r"^MUFU"),
    ("FP64 (DADD / DMUL / DFMA / DSETP)", r"^D(ADD|MUL|FMA|SETP)"), ("tensor core (HMMA / UTCMMA / ...)", r"^(HMMA|IMMA|DMMA|UTC)"),
]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), [])
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)(.*?);", line)
        if m and cur is not None:
            cur.append(m.group(1) + m.group(2))
    demangle = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# SASS opcode summary of %s (sm_100a), %d kernels" % (os.path.relpath(LIB, ROOT), len(kernels)))
    total = collections.Counter()
    for (name, ins), dn in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", dn)
        if "rl::" not in dn:  # CUB kernels (sort / scan used by the table builds) are library code
            continue
        print("\n## %s   [%d instructions]" % (short[:140], len(ins)))
        for label, pat in GROUPS:
            n = sum(1 for i in ins if re.search(pat, i))
            total[label] += n
            if n:
                print("  %-62s %5d" % (label, n))
    print("\n## all rl:: kernels")
    for label, _ in GROUPS:
        print("  %-62s %6d" % (label, total[label]))


if __name__ == "__main__":
    main()
