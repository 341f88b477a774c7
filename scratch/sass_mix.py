"""Opcode mix (executed warp instructions) and stall samples per opcode class from an ncu source page.
usage: python scratch/sass_mix.py report.ncu-rep [kernel-substring]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name",')
for b in blocks[1:]:
    lines = b.splitlines()
    name = lines[0].strip('",')
    if want not in name: continue
    rd = csv.DictReader(io.StringIO("\n".join(lines[1:])))
    mix = collections.Counter(); stall = collections.Counter(); tot = 0; stot = 0
    for r in rd:
        try: n = int(r["Instructions Executed"]); s = int(r["# Samples"])
        except Exception: continue
        src = r["Source"].strip()
        if src.startswith("@"): src = src.split(None, 1)[1]
        op = src.split()[0].split(".")[0]
        mix[op] += n; stall[op] += s; tot += n; stot += s
    print(name, "total warp inst", tot, "samples", stot)
    for op, n in mix.most_common(28):
        print("  %-10s %6.2f %%   samples %5.2f %%" % (op, 100.0 * n / tot, 100.0 * stall[op] / max(stot, 1)))
