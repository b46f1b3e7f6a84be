"""Join an ncu SASS source-page CSV with `nvdisasm -g -c` output to get executed
warp-instructions per CUDA source line.
usage: linemix.py src.csv disasm.txt kernel_substring n_warps source_file.cu"""
import csv, re, sys, collections
srccsv, dis, kname, nw, cu = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4]), sys.argv[5]
rows = list(csv.reader(open(srccsv)))
hdr = next(r for r in rows if 'Source' in r and 'Instructions Executed' in r)
iE, iW = hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
ex = [(int(r[iE]), int(r[iW] or 0)) for r in rows if len(r) > iE and r[iE].isdigit()]
lines, cur, on = [], None, False
for l in open(dis):
    if l.startswith('//---') and '.text.' in l:
        on = kname in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if m.group(1).endswith(cu.split('/')[-1]):
            cur = int(m.group(2))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        lines.append(cur)
print('sass instrs: ncu', len(ex), 'nvdisasm', len(lines))
agg, st = collections.Counter(), collections.Counter()
for (e, w), ln in zip(ex, lines):
    agg[ln] += e; st[ln] += w
src = open(cu).read().split('\n')
allw = sum(st.values()) or 1
for ln in sorted(agg):
    if agg[ln]/nw >= 2 or st[ln]/allw > 0.01:
        print(f'{agg[ln]/nw:7.1f} {100*st[ln]/allw:5.1f}%  {ln:4d}: {src[ln-1].strip()[:100] if ln else ""}')
