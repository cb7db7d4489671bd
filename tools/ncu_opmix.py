"""Per-opcode and per-source-line executed-instruction mix of one kernel from an .ncu-rep
(needs the SourceCounters section). Usage: python tools/ncu_opmix.py REPORT [warp_iterations]"""
import collections
import csv
import io
import re
import subprocess
import sys


def page(report, *extra):
    out = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv", *extra], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    report = sys.argv[1]
    denom = float(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = page(report, "--print-source", "sass")
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    h = rows[starts[0]]
    isrc, iex = h.index("Source"), h.index("Instructions Executed")
    end = starts[1] - 1 if len(starts) > 1 else len(rows)
    data = [r for r in rows[starts[0] + 1:end] if len(r) > iex and r[iex].isdigit()]
    total = sum(int(r[iex]) for r in data)
    print("kernel:", rows[0][1] if rows and len(rows[0]) > 1 else "?", " executed warp instructions:", total,
          (" per unit: %.1f" % (total / denom)) if denom else "")
    ops = collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)", r[isrc])
        ops[m.group(2) if m else "?"] += int(r[iex])
    for k, v in ops.most_common(32):
        print("%-8s %6.2f%%%s" % (k, 100.0 * v / total, ("  %7.2f/unit" % (v / denom)) if denom else ""))


if __name__ == "__main__":
    main()
