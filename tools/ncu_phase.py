"""Per-phase (source line range) samples, instructions and stall reasons from an .ncu-rep.

usage: ncu_phase.py report.ncu-rep file.cu line1 line2 ...   (boundaries in file.cu)
       add --lines N to also list the N hottest source lines with their top stall reasons
"""
import csv, subprocess, sys, io, collections
args = [a for a in sys.argv[1:]]
top = 0
if "--lines" in args:
    i = args.index("--lines"); top = int(args[i + 1]); del args[i:i + 2]
rep, fname = args[0], args[1]
bounds = [int(x) for x in args[2:]]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
STALLS = ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_no_inst", "stall_mio", "stall_math",
          "stall_branch_resolving", "stall_not_selected", "stall_selected", "stall_dispatch", "stall_lg",
          "stall_barrier"]
agg = collections.defaultdict(lambda: collections.Counter())
lines = collections.defaultdict(lambda: collections.Counter())
text = {}
cur_file = None; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; idx = {k: hdr.index(k) for k in STALLS + ['# Samples', 'Instructions Executed']}; continue
    if hdr is None or len(r) < len(hdr) or r[2] != '-': continue
    try: smp = int(r[idx['# Samples']]); ins = int(r[idx['Instructions Executed']])
    except ValueError: continue
    if cur_file == fname:
        ln = int(r[0]); key = fname + ':' + str(max([b for b in bounds if b <= ln] or [0]))
    else: key = cur_file
    for tgt in (agg[key], lines[(cur_file, int(r[0]))]):
        tgt['smp'] += smp; tgt['ins'] += ins
        for s in STALLS:
            try: tgt[s] += int(r[idx[s]])
            except ValueError: pass
    text[(cur_file, int(r[0]))] = r[1]
ts = sum(v['smp'] for v in agg.values()); ti = sum(v['ins'] for v in agg.values())
print(f"total samples {ts}, total warp instructions {ti}")
print(f"{'range':28s} {'smp%':>6s} {'ins%':>6s}  " + " ".join(f"{s[6:12]:>6s}" for s in STALLS))
for k, v in sorted(agg.items()):
    print(f"{k:28s} {100*v['smp']/ts:6.1f} {100*v['ins']/ti:6.1f}  " +
          " ".join(f"{100*v[s]/max(ts,1):6.1f}" for s in STALLS))
tot = collections.Counter()
for v in agg.values(): tot.update(v)
print(f"{'ALL':28s} {100.0:6.1f} {100.0:6.1f}  " + " ".join(f"{100*tot[s]/max(ts,1):6.1f}" for s in STALLS))
if top:
    for k, v in sorted(lines.items(), key=lambda kv: -kv[1]['smp'])[:top]:
        st = sorted(STALLS, key=lambda s: -v[s])[:3]
        print(f"{100*v['smp']/ts:5.1f}% smp {100*v['ins']/ti:5.1f}% ins {k[0]}:{k[1]:<5d} "
              + ",".join(f"{s[6:]}={100*v[s]/max(v['smp'],1):.0f}%" for s in st) + "  " + text[k].strip()[:80])
