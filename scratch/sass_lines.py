"""Executed warp instructions per full opcode (with modifiers) / per source line.
usage: python scratch/sass_lines.py report.ncu-rep kernel-substring [op-prefix-for-line-listing]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; want = sys.argv[2]; opf = sys.argv[3] if len(sys.argv) > 3 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda" if False else "sass"], capture_output=True, text=True).stdout
for b in out.split('"Kernel Name",')[1:]:
    lines = b.splitlines(); name = lines[0].strip('",')
    if want not in name: continue
    rd = csv.DictReader(io.StringIO("\n".join(lines[1:])))
    mix = collections.Counter(); tot = 0
    for r in rd:
        try: n = int(r["Instructions Executed"])
        except Exception: continue
        src = r["Source"].strip()
        if src.startswith("@"): src = src.split(None, 1)[1]
        op = src.split()[0]
        mix[op] += n; tot += n
    print(name, tot)
    for op, n in mix.most_common(40):
        if opf is None or op.startswith(opf): print("  %-22s %6.2f %%" % (op, 100.0 * n / tot))
    break
