"""Per-SASS-instruction executed counts and stall samples of one kernel from an .ncu-rep (source page)."""
import csv
import subprocess
import sys

rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', kernel], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# several launches may follow each other: take the first
hdr = None
data = []
n_k = 0
for r in rows:
    if r and r[0] == 'Kernel Name':
        n_k += 1
        if n_k > 1:
            break
        continue
    if r and r[0] == 'Address':
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(r)
ix = hdr.index('Instructions Executed')
isamp = hdr.index('# Samples')
isrc = hdr.index('Source')
tot = sum(int(r[ix]) for r in data)
tots = sum(int(r[isamp]) for r in data)
print('total warp-instructions %d, samples %d, sass lines %d' % (tot, tots, len(data)))
if top:
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:top]:
        print('%8d %6.2f%%  samples %6d %5.1f%%  %s' % (int(r[ix]), 100.0 * int(r[ix]) / tot, int(r[isamp]), 100.0 * int(r[isamp]) / max(tots, 1), r[isrc].strip()))
else:
    for i, r in enumerate(data):
        print('%4d %9d %5.2f%% s%6d %5.1f%%  %s' % (i, int(r[ix]), 100.0 * int(r[ix]) / tot, int(r[isamp]), 100.0 * int(r[isamp]) / max(tots, 1), r[isrc].strip()))
