"""Hottest CUDA source lines (samples, instructions executed) from an .ncu-rep."""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
agg = {}
cur_file = None
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[2] != '-': continue   # only per-source-line aggregate rows (Address == '-')
    try:
        smp = int(r[hdr.index('# Samples')]); ins = int(r[hdr.index('Instructions Executed')])
    except ValueError:
        continue
    key = (cur_file, int(r[0]))
    a = agg.setdefault(key, [0, 0, r[1]])
    a[0] += smp; a[1] += ins
tot = sum(a[0] for a in agg.values()); toti = sum(a[1] for a in agg.values())
print('total samples', tot, 'total warp-inst', toti)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*a[0]/tot:5.1f}% smp {100*a[1]/toti:5.1f}% ins {k[0]}:{k[1]:<5d} {a[2][:100]}")
