"""Phase statistics of a tools/iter_trace.py timeline (math warp 4 of CTA 0): where a sub-tile's time goes."""
import re
import sys
from collections import defaultdict

rows = []
for line in open(sys.argv[1]):
  m = re.match(r'\s*([\d.]+) us\s+(\S+)\s+(\d+)', line)
  if m:
    rows.append((float(m.group(1)), m.group(2), int(m.group(3))))
math_kinds = ('E_BEGIN', 'E_SUB', 'e_in', 'e_ld', 'e_cmp', 'e_yw', 'e_arr', 'E_END')
seq = [r for r in rows if r[1] in math_kinds]
dur = defaultdict(list)
for (t0, k0, _), (t1, k1, _) in zip(seq, seq[1:]):
  dur['%s->%s' % (k0, k1)].append(t1 - t0)
span = seq[-1][0] - seq[0][0]
print('math warp timeline: %.1f us, %d events' % (span, len(seq)))
for k, v in sorted(dur.items(), key=lambda kv: -sum(kv[1])):
  print('  %-18s n=%4d  mean %.3f us  total %.1f us (%.0f%%)' % (k, len(v), sum(v) / len(v), sum(v), 100 * sum(v) / span))
g = [r for r in rows if r[1] in ('G_BEGIN', 'G_END')]
gb = {i: t for t, k, i in g if k == 'G_BEGIN'}
ge = {i: t for t, k, i in g if k == 'G_END'}
d = [ge[i] - gb[i] for i in gb if i in ge]
if d:
  print('G tiles: %d, mean duration %.2f us, period %.2f us' % (len(d), sum(d) / len(d), (max(ge.values()) - min(gb.values())) / len(d)))
gl = [t for t, k, i in rows if k == 'G_LOAD']
if len(gl) > 1:
  print('G_LOAD: %d, mean interval %.3f us' % (len(gl), (gl[-1] - gl[0]) / (len(gl) - 1)))
