"""Print selected metrics of every launch in an .ncu-rep (reads the raw page; no GPU needed)."""
import csv, re, subprocess, sys
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else '.'
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('==', d['Kernel Name'][:90], 'grid', d.get('Grid Size'), 'block', d.get('Block Size'))
    for h, u in zip(hdr, units):
        if re.search(pat, h):
            print('  %-95s %-10s %s' % (h, u, d[h]))
