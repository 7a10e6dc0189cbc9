"""Per-source-line instruction counts from an ncu report captured with --import-source on.

    python profiles/tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]

Reads `ncu -i REP --page source --csv --print-source cuda,sass` (needs -lineinfo at compile time)
and prints, per kernel and source file, the lines with the most executed warp instructions:
instructions (millions), average active threads per instruction, stall samples, line, source.
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    fname = func = None
    col = None
    per = defaultdict(lambda: defaultdict(list))       # func -> file -> [(inst, thr, smp, line, src)]
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            func = r[1].split("(")[0].replace("void ", "").replace("pmr::", "")
            continue
        if r[0] == "Line No":
            col = {}
            for i, c in enumerate(r):
                col.setdefault(c, i)
            continue
        if col is None or not r[0].strip().isdigit():
            continue
        try:
            inst = float(r[col["Instructions Executed"]])
            thr = float(r[col["Thread Instructions Executed"]])
            smp = float(r[col["# Samples"]])
        except (ValueError, KeyError, IndexError):
            continue
        if inst > 0:
            per[func][fname].append((inst, thr, smp, int(r[0]), r[1].strip()))
    for func, files in per.items():
        total = sum(e[0] for v in files.values() for e in v)
        print("######## %s: %.1fM warp instructions" % (func, total / 1e6))
        for fname, v in sorted(files.items(), key=lambda kv: -sum(e[0] for e in kv[1])):
            print("===== %s inst %.1fM samples %d" % (fname, sum(e[0] for e in v) / 1e6, sum(e[2] for e in v)))
            for inst, thr, smp, line, src in sorted(v, key=lambda e: -e[0])[:top]:
                print("  %6.2fM thr %4.1f smp %5d  %4d | %s" % (inst / 1e6, thr / inst, smp, line, src[:120]))


if __name__ == "__main__":
    main()
