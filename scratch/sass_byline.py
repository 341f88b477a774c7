"""Per CUDA source line: executed warp instructions by class, from `ncu --page source --print-source cuda,sass --csv`.
usage: python scratch/sass_byline.py report.ncu-rep kernel-substring [opcode-prefix]   (prints top lines for that opcode, or overall)"""
import csv, subprocess, sys, collections, io
rep, want = sys.argv[1], sys.argv[2]
opf = sys.argv[3] if len(sys.argv) > 3 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None; cur_fn = None; hdr = None; line = None; text = None
per = collections.Counter(); tot = 0; src = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": cur_fn = r[1]; continue
    if r[0] == "Line No": hdr = r; iE = hdr.index("Instructions Executed"); continue
    if hdr is None or want not in (cur_fn or ""): continue
    if r[0] != "-" and r[0] != "":
        line = (cur_file, int(r[0])); src[line] = r[1].strip(); 
        if r[2] == "-": continue
    sass = r[3].strip()
    if not sass or sass == "-": continue
    try: n = int(r[iE])
    except Exception: continue
    if sass.startswith("@"): sass = sass.split(None, 1)[1]
    op = sass.split()[0]
    tot += n
    if opf is None or op.startswith(opf): per[line] += n
print("total", tot, "selected", sum(per.values()), "= %.2f %%" % (100.0 * sum(per.values()) / tot))
for l, n in per.most_common(40):
    print("%6.2f %%  %s:%d  %s" % (100.0 * n / tot, l[0], l[1], src.get(l, "")[:110]))
