"""Groups the SASS of one kernel of an .ncu-rep (captured with --import-source on) into regions of equal execution
count -- loop bodies, per-round and per-batch code -- with their share of the executed instructions and of the samples.

    python tools/ncu_source_regions.py <report.ncu-rep> <kernel-regex> > profiles/<name>.txt
"""
import csv
import io
import subprocess
import sys
from collections import Counter


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = rows[1]
    isrc, iex, ism = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ins = [(r[isrc].strip(), int(r[iex]), int(r[ism])) for r in rows[2:] if len(r) > iex and r[iex].isdigit()]
    if len(ins) % 2 == 0 and [x[0] for x in ins[:len(ins) // 2]] == [x[0] for x in ins[len(ins) // 2:]]:
        ins = ins[:len(ins) // 2]  # the page lists the kernel twice
    tot, tsm = sum(x[1] for x in ins), sum(x[2] for x in ins)
    print("# %s, kernel %s: %d warp instructions, %d samples" % (rep, rows[0][1][:90] if len(rows[0]) > 1 else rx, tot, tsm))
    groups = []
    for k, x in enumerate(ins):
        op = x[0].split()[1] if x[0].startswith("@") else x[0].split()[0]
        if groups and abs(x[1] - groups[-1]["e"]) <= 0.03 * max(groups[-1]["e"], 1):
            g = groups[-1]
            g["n"] += 1; g["sum"] += x[1]; g["samp"] += x[2]; g["ops"].append(op); g["end"] = k
        else:
            groups.append({"start": k, "end": k, "e": x[1], "n": 1, "sum": x[1], "samp": x[2], "ops": [op]})
    print("%-11s %5s %12s %8s %9s  %s" % ("sass lines", "n", "exec/inst", "share", "samples", "opcodes"))
    for g in groups:
        if g["sum"] / max(tot, 1) > 0.004:
            c = Counter(g["ops"])
            print("%4d-%-6d %5d %12.3g %7.2f%% %8.2f%%  %s" % (g["start"], g["end"], g["n"], g["e"], 100 * g["sum"] / tot,
                                                              100 * g["samp"] / max(tsm, 1),
                                                              " ".join("%s:%d" % kv for kv in c.most_common(7))))


if __name__ == "__main__":
    main()
