"""Code size per source region from nvdisasm --print-line-info output of the fast kernel."""
import re, sys, collections, subprocess, os, glob, tempfile
so = sys.argv[1] if len(sys.argv) > 1 else 'sbdart_b200/libsbdart_b200.so'
key = sys.argv[2] if len(sys.argv) > 2 else 'disort_fast_kernelILi8'
d = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=d, capture_output=True)
for cub in glob.glob(d + '/*.cubin'):
    txt = subprocess.run(['nvdisasm', '--print-line-info', cub], capture_output=True, text=True).stdout
    lines = txt.splitlines()
    idx = [i for i, l in enumerate(lines) if l.startswith('.text.') and key in l]
    if not idx: continue
    i0 = idx[0]
    i1 = next((i for i in range(i0 + 1, len(lines)) if lines[i].startswith('.text.')), len(lines))
    cur = None; cnt = collections.Counter()
    for l in lines[i0:i1]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
        if re.match(r'\s+/\*[0-9a-f]{4,6}\*/', l): cnt[cur] += 1
    print('total', sum(cnt.values()))
    b = collections.Counter()
    for (fn, ln), v in cnt.items():
        if fn in ('sbd_fast.cu', 'sbd_adding.cu', 'sbd_wide.cu'): b[(ln // 20) * 20] += v
        else: b[fn] += v
    for k, v in sorted(b.items(), key=lambda kv: -kv[1])[:28]: print(v, k)
    break
