#!/usr/bin/env python
"""FP64-pipe thread instructions (DADD / DMUL / DFMA / DSETP) per cell of every kernel in an ncu
report that was captured with --import-source on.

    python profiles/fp64_count.py gpurun_out/prof_X.ncu-rep cells_per_launch
"""
import csv
import io
import subprocess
import sys

rep, cells = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
for block in out.split('"Kernel Name",')[1:]:
    lines = block.splitlines()
    name = lines[0].strip('",')
    fp64 = total = 0
    for r in csv.DictReader(io.StringIO("\n".join(lines[1:]))):
        try:
            n = int(r["Thread Instructions Executed"])
        except (ValueError, KeyError, TypeError):
            continue
        src = r["Source"].strip()
        if src.startswith("@"):
            src = src.split(None, 1)[1]
        op = src.split()[0].split(".")[0]
        total += n
        if op in ("DADD", "DMUL", "DFMA", "DSETP"):
            fp64 += n
    print("%-90s fp64/cell %.1f  all/cell %.1f" % (name[:90], fp64 / cells, total / cells))
