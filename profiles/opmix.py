"""Aggregate an `ncu --page source --csv` dump by SASS opcode.
usage: python profiles/opmix.py src.csv n_warps"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
nw = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = next(r for r in rows if 'Source' in r and 'Instructions Executed' in r)
iS, iE, iW = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
ops, stall, tot, n = collections.Counter(), collections.Counter(), 0, 0
for r in rows:
    if r is hdr or len(r) <= iE or not r[iE].isdigit():
        continue
    parts = r[iS].split()
    if not parts:
        continue
    op = parts[1] if parts[0].startswith('@') else parts[0]
    key = op if op.startswith(('LDG', 'STG', 'F2I', 'I2F', 'MUFU', 'DSETP', 'FRND')) else op.split('.')[0]
    e = int(r[iE]); ops[key] += e; tot += e; stall[key] += int(r[iW] or 0); n += 1
print(f'static instrs {n}, dynamic warp instrs {tot}, per warp {tot/nw:.1f}')
allst = sum(stall.values()) or 1
for op, c in ops.most_common(50):
    print(f'{op:28s} {c/nw:8.1f} /warp   stall {100*stall[op]/allst:5.1f}%')
