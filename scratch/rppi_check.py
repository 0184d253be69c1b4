"""Scratch: tiled (r_p, Pi) kernel against the general kernel (and the oracle for small cases) over many configurations."""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
from measure_ia_b200 import MeasureIABox
from measure_ia_b200.synthetic import uniform_box

cases = [
	# n, L, seed, jk, n_r, n_mu, kwargs
	(3000, 205.0, 1, 0, 10, 8, {}),
	(3000, 205.0, 1, 27, 10, 8, {}),
	(20000, 205.0, 2, 27, 10, 8, dict(weights=True)),
	(20000, 205.0, 3, 27, 8, 20, dict(los=0)),
	(20000, 150.0, 4, 8, 6, 12, dict(los=1, weights=True)),
	(5000, 50.0, 5, 8, 4, 4, {}),          # tiny box: all-mode, straddles
	(2000, 30.0, 6, 27, 5, 5, {}),         # r_max > L/2
	(50, 50.0, 7, 8, 4, 4, {}),
	(33, 50.0, 8, 8, 4, 4, {}),
	(1, 50.0, 9, 8, 4, 4, {}),
	(60000, 300.0, 10, 64, 8, 20, dict(n_shape=30000, weights=True, clustered=0.5)),
	(100000, 205.0, 11, 27, 10, 8, {}),
	(100000, 205.0, 12, 27, 10, 1, {}),
	(100000, 205.0, 13, 125, 10, 10, dict(los=1)),
	(30000, 205.0, 14, 27, 10, 8, dict(periodicity=False)),
	(50000, 205.0, 15, 27, 10, 8, dict(pi_max=30.0)),
	(50000, 205.0, 16, 27, 8, 20, dict(pi_max=60.0, los=1, weights=True)),
	(200000, 205.0, 17, 27, 10, 8, {}),
]
only = [int(x) for x in sys.argv[1:]]
bad = 0
for ci, (n, L, seed, jk, n_r, n_mu, kw) in enumerate(cases):
	if only and ci not in only:
		continue
	kw = dict(kw)
	per = kw.pop("periodicity", True)
	pimax = kw.pop("pi_max", None)
	d = uniform_box(n, L, seed=seed, **kw)
	res = {}
	for kern in ('general', 'tiled'):
		b = MeasureIABox(d, None, boxsize=L, num_bins_r=n_r, num_bins_pi=n_mu, periodicity=per, pi_max=pimax)
		b.kernel = kern
		t0 = time.perf_counter()
		try:
			b.measure_xi_w('a', 'both', jk, temp_file_path=False)
		except Exception as e:
			print(ci, kern, 'FAILED', repr(e)[:300])
			bad += 1
			res[kern] = None
			continue
		dt = time.perf_counter() - t0
		res[kern] = (b.last_result, b.last_stats)
		st = b.last_stats
		print(ci, kern, 'n', n, 'L', L, 'jk', jk, 'bins', n_r, n_mu, kw, 'tested', st['tested'], 'binned', st['binned'], 'nan', st['nan_rule'],
			  'tasks', st['tasks'], 'kernel', st['kernel'], 'pairs_ms', round(st['phases_ms']['pairs'], 3), 'wall', round(dt, 3))
	if res.get('general') is None or res.get('tiled') is None:
		continue
	g, t = res['general'][0], res['tiled'][0]
	ok = np.array_equal(g['count'], t['count']) and np.array_equal(g['count_jk'], t['count_jk'])
	msg = []
	for k in ('DD', 'SpD_raw', 'ScD_raw', 'DD_jk', 'SpD_jk'):
		a, c = np.asarray(g[k]), np.asarray(t[k])
		if a.size == 0:
			continue
		tol = 1e-10 * np.abs(a) + 1e-11 * np.abs(a).max()
		e = np.abs(a - c)
		if not (e <= tol).all():
			ok = False
		msg.append(f"{k}:{(e / (np.abs(a).max() + 1e-300)).max():.1e}")
	print(ci, 'MATCH' if ok else 'MISMATCH', 'count diff', int(np.abs(g['count'] - t['count']).sum()), ' '.join(msg))
	if not ok:
		bad += 1
		print((t['count'] - g['count']))
print('BAD', bad)
