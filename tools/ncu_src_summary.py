"""Summarise an `ncu --page source --csv` dump: opcode mix, stall reasons, hottest SASS lines."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr) and r[0] != 'Address']
if '--first' in sys.argv:
    n = next((i for i, r in enumerate(rows[2:]) if r and r[0] == 'Kernel Name'), None)
    data = [r for r in rows[2:(n + 2 if n is not None else None)] if len(r) == len(hdr) and r[0] != 'Address']
ix = {h: i for i, h in enumerate(hdr)}
num = lambda r, k: int(r[ix[k]] or 0)
tot_inst = sum(num(r, 'Instructions Executed') for r in data)
tot_samp = sum(num(r, '# Samples') for r in data)
print('total warp instrs', tot_inst, 'samples', tot_samp, 'n sass', len(data))
ops, samp = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ix['Source']])
    op = (m.group(2) if m else r[ix['Source']][:20]).split('.')[0]
    ops[op] += num(r, 'Instructions Executed')
    samp[op] += num(r, '# Samples')
for op, c in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"{op:12s} {c:12d} {100*c/tot_inst:5.1f}%  samples {samp[op]:7d} {100*samp[op]/max(1,tot_samp):5.1f}%")
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = {s: sum(num(r, s) for r in data) for s in stalls}
print(sorted(tot.items(), key=lambda kv: -kv[1])[:10])
for r in sorted(data, key=lambda r: -num(r, '# Samples'))[:30]:
    print(r[ix['# Samples']], r[ix['Instructions Executed']], r[ix['Source']][:100],
          {s[6:]: num(r, s) for s in stalls if num(r, s) > 0.1 * max(1, num(r, '# Samples'))})
