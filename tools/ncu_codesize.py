import csv, subprocess, sys, io, collections
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cnt = collections.Counter(); text = {}
cur_file=None; hdr=None; cur_line=None
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur_file=r[1].split('/')[-1]; continue
    if r[0]=="Line No": hdr=r; continue
    if hdr is None or len(r)<4: continue
    if r[2]=='-':
        cur_line=(cur_file,int(r[0])); text[cur_line]=r[1]
    else:
        cnt[cur_line]+=1
tot=sum(cnt.values()); print('total sass', tot)
# aggregate in ranges of source lines
for k,v in cnt.most_common(int(sys.argv[2]) if len(sys.argv)>2 else 30):
    print(v, k, text[k][:90])
