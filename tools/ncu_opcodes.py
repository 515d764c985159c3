"""Executed warp-instructions per SASS opcode (and per source-line range) from an .ncu-rep.
usage: ncu_opcodes.py report.ncu-rep [file.cu line1 line2 ...]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
fname = sys.argv[2] if len(sys.argv) > 2 else None
bounds = [int(x) for x in sys.argv[3:]]
# The SASS-only view lists every instruction once (exact totals); the cuda,sass view
# repeats an instruction under every frame of its inline chain and is only used for
# the per-range split (shares there are approximate).
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass" if fname else "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if not fname:
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    ie = rows[hi].index('Instructions Executed'); isrc = rows[hi].index('Source')
    tot = collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) > ie and r[ie].isdigit():
            t = r[isrc].split()
            op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
            tot[op] += int(r[ie])
    T = sum(tot.values()); print("total warp instructions", T)
    for op, c in tot.most_common(25): print(f"{op:10s} {c:12d} {100*c/T:5.1f}%")
    sys.exit(0)
tot = collections.Counter(); byphase = collections.defaultdict(collections.Counter)
cur = None; hdr = None; key = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; ie = hdr.index('Instructions Executed'); continue
    if hdr is None or len(r) < len(hdr): continue
    if r[2] == '-':
        ln = int(r[0])
        key = (cur + ':' + str(max([b for b in bounds if b <= ln] or [0]))) if cur == fname else cur
        continue
    try: ex = int(r[ie])
    except ValueError: continue
    op = r[3].split()[0] if not r[3].startswith('@') else r[3].split()[1]
    op = op.split('.')[0]
    tot[op] += ex; byphase[key][op] += ex
T = sum(tot.values())
print("total", T)
for op, c in tot.most_common(25): print(f"{op:10s} {c:12d} {100*c/T:5.1f}%")
if fname:
    for k in sorted(byphase):
        c = byphase[k]; t = sum(c.values())
        print(f"{k:28s} {100*t/T:5.1f}%  " + " ".join(f"{o}={100*v/T:.1f}" for o, v in c.most_common(6)))
