"""Top stalled SASS instructions of launch #idx in an .ncu-rep (source page)."""
import csv, subprocess, sys
rep = sys.argv[1]; idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
lo = int(sys.argv[4], 16) if len(sys.argv) > 4 else None; hi = int(sys.argv[5], 16) if len(sys.argv) > 5 else None
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
starts.append(len(rows))
blk = rows[starts[idx]:starts[idx + 1]]
print(blk[0][1][:120])
hdr = blk[1]; ix = {k: i for i, k in enumerate(hdr)}
data = [r for r in blk[2:] if len(r) == len(hdr)]
tot = sum(int(r[ix['# Samples']]) for r in data)
print('total samples', tot)
stalls = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
agg = {k: sum(int(r[ix[k]]) for r in data) for k in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
if lo is not None:
    for r in data:
        a = int(r[ix['Address']], 16) & 0xfffff
        if lo <= a <= hi:
            print('%05x %6s %9s  %s' % (a, r[ix['# Samples']], r[ix['Instructions Executed']], r[ix['Source']].strip()[:90]))
else:
    for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:top]:
        s = {k: int(r[ix[k]]) for k in stalls if int(r[ix[k]]) > 0}
        s = dict(sorted(s.items(), key=lambda kv: -kv[1])[:2])
        print(r[ix['Address']][-5:], '%6d' % int(r[ix['# Samples']]), '%9d' % int(r[ix['Instructions Executed']]), r[ix['Source']].strip()[:64], s)
