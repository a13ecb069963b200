"""Pretty-print a fit trace (tools/fit_trace.py): per block column the chain kernels, and per stream the busy time."""
import sys, json, collections
rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith('{')]
print(rows[0])
rows = rows[1:]
by = collections.defaultdict(dict)
for r in rows:
    by[r['kb']][r['k']] = (r['t0_us'], r['t1_us'])
names = ['chol', 'trsm', 'colupd', 'inv128', 'inv_row', 'inv_upd', 'inv_new', 'bulk', 'w_acc']
print('kb   ' + ''.join('%-19s' % n for n in names))
for kb in sorted(by):
    print('%-4d ' % kb + ''.join(('%7.0f-%-5.0f(%3.0f) ' % (by[kb][n][0], by[kb][n][1], by[kb][n][1] - by[kb][n][0])) if n in by[kb] else ' ' * 19 for n in names))
dur = collections.defaultdict(float)
for r in rows:
    dur[r['k']] += r['t1_us'] - r['t0_us']
print({k: round(v) for k, v in dur.items()})
