import csv, subprocess, sys, io, collections
rep = sys.argv[1]
bounds = [int(x) for x in sys.argv[2:]]   # line boundaries in sbd_fast.cu
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
agg = collections.defaultdict(lambda: [0, 0])
cur_file = None; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != '-': continue
    try: smp = int(r[hdr.index('# Samples')]); ins = int(r[hdr.index('Instructions Executed')])
    except ValueError: continue
    if cur_file == 'sbd_fast.cu':
        ln = int(r[0]); key = 'fast:' + str(max([b for b in bounds if b <= ln] or [0]))
    else: key = cur_file
    agg[key][0] += smp; agg[key][1] += ins
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items()): print(f"{k:28s} {100*v[0]/ts:5.1f}% samples {100*v[1]/ti:5.1f}% instr")
