"""CPU: the two facts the light-cone kernel (measure_ia_b200/csrc/mia_lightcone.cuh) relies on, checked in numpy with the
kernel's own operation sequences.

1. The shape projection needs no transcendental per pair: with E1 = e cos 2phi_axis, E2 = e sin 2phi_axis,
       e_+ = -(E1 c2 + E2 s2),  e_x = -(E2 c2 - E1 s2),  c2 = (dx^2 - dy^2) / r_p^2,  s2 = 2 dx dy / r_p^2
   equals the reference's  -e cos 2(phi_axis - arctan2(dy / r_p, dx / r_p)),  -e sin 2(...)  (measure_w_lightcone.py:148-152)
   to a few ulp of e -- far inside the 1e-10 contract on the sums.
2. The sky pre-filter never drops a pair the reference would bin: if |ddec| (or |dra|) exceeds reach / (pi / 180 chi scale [cos dec]),
   reach = sqrt(last r threshold) (1 + 1e-9), then the reference's exactly rounded r_p^2 (resp. r^2) is >= that threshold."""
import numpy as np


def _exact_offsets(ra_s, dec_s, ra_n, dec_n, chi_n, scale):
	"""measure_w_lightcone.py:140-147 for one position galaxy against arrays of shape galaxies."""
	dra = (ra_s - ra_n) / 180 * np.pi
	ddec = (dec_s - dec_n) / 180 * np.pi
	dx = dra * chi_n * np.cos(dec_n / 180 * np.pi)
	dy = ddec * chi_n
	return dx * scale, dy * scale


def test_projection_without_transcendentals():
	rng = np.random.default_rng(2)
	n = 200_000
	dx, dy = rng.normal(0, 5, n), rng.normal(0, 5, n)
	dx[:1000] *= 1e-6  # nearly vertical / horizontal separations too
	dy[1000:2000] *= 1e-6
	e1, e2 = rng.normal(0, 0.3, n), rng.normal(0, 0.3, n)
	# the reference's chain (measure_w_lightcone.py:135-152)
	theta = 1. / 2 * np.arctan2(e2, e1)
	axis = np.array([np.cos(theta), np.sin(theta)])
	axis = axis / np.sqrt(np.sum(axis ** 2, axis=0))
	e = np.sqrt(e1 ** 2 + e2 ** 2)
	phi_axis = np.arctan2(axis[1], axis[0])
	rp = np.sqrt(dx ** 2 + dy ** 2)
	phi = phi_axis - np.arctan2(dy / rp, dx / rp)
	want_p, want_c = -e * np.cos(2 * phi), -e * np.sin(2 * phi)
	# the kernel's sequence
	E1, E2 = e * np.cos(2 * phi_axis), e * np.sin(2 * phi_axis)
	inv = 1.0 / (dx * dx + dy * dy)
	c2, s2 = (dx * dx - dy * dy) * inv, 2.0 * dx * dy * inv
	got_p, got_c = -(E1 * c2 + E2 * s2), -(E2 * c2 - E1 * s2)
	assert np.max(np.abs(got_p - want_p) / e) < 2e-15 and np.max(np.abs(got_c - want_c) / e) < 2e-15


def test_sky_prefilter_is_conservative():
	rng = np.random.default_rng(3)
	r_max2 = 20.000000000000004 ** 2  # the last threshold on r_p^2 for the default limits
	reach = np.sqrt(r_max2) * (1.0 + 1e-9)
	for scale in (1.0, 0.7):
		for _ in range(200):
			ra_n, dec_n, chi_n = rng.uniform(0, 360), rng.uniform(-80, 80), rng.uniform(50, 3000)
			k_dec = (np.pi / 180.0) * chi_n * scale
			lim_dec, lim_ra = reach / k_dec, reach / (k_dec * abs(np.cos(dec_n / 180 * np.pi)))
			# shape galaxies right at the limits, a hair inside and outside (where a sloppy filter would go wrong)
			f = 1.0 + rng.uniform(-3e-9, 3e-9, 4000)
			sgn = rng.choice([-1.0, 1.0], 4000)
			dec_s = dec_n + sgn * lim_dec * f
			ra_s = ra_n + rng.uniform(-0.5, 0.5, 4000) * lim_ra
			px, py = _exact_offsets(ra_s, dec_s, ra_n, dec_n, chi_n, scale)
			dropped = np.abs(dec_s - dec_n) > lim_dec
			assert dropped.any() and (~dropped).any()
			assert np.all((px ** 2 + py ** 2)[dropped] >= r_max2)
			ra_s2 = ra_n + sgn * lim_ra * f
			dec_s2 = dec_n + rng.uniform(-0.5, 0.5, 4000) * lim_dec
			px, py = _exact_offsets(ra_s2, dec_s2, ra_n, dec_n, chi_n, scale)
			dropped = np.abs(ra_s2 - ra_n) > lim_ra
			assert np.all((px ** 2 + py ** 2)[dropped] >= r_max2)


def test_chi_window_is_a_superset():
	"""The per-CTA window [chi(first) + win_lo - slack, chi(last) + win_hi + slack] (mia_lightcone.cuh) holds every shape galaxy whose
	exactly rounded Pi = chi_s - chi_n passes the reference's range mask for some position galaxy of the block."""
	rng = np.random.default_rng(4)
	lo_thr, hi_thr = -60.0, 60.0
	for _ in range(300):
		chi_p = np.sort(rng.uniform(100.0, 3000.0) + rng.uniform(0, 5.0, 128))
		c0, c1 = chi_p[0], chi_p[-1]
		lo = c0 + lo_thr - 1e-9 * (abs(c0) + abs(lo_thr)) - 1e-300
		hi = c1 + hi_thr + 1e-9 * (abs(c1) + abs(hi_thr)) + 1e-300
		# shape galaxies within a few ulp of the window's ends, as seen from the first / last position galaxy
		eps = np.arange(-8, 9)
		chi_s = np.concatenate([np.nextafter(c0 + lo_thr, np.inf) + eps * np.spacing(c0), np.nextafter(c1 + hi_thr, -np.inf) + eps * np.spacing(c1)])
		pi = chi_s[None, :] - chi_p[:, None]
		binned_by_some = np.any((pi >= lo_thr) & (pi < hi_thr), axis=0)
		inside = (chi_s >= lo) & (chi_s <= hi)
		assert np.all(inside[binned_by_some])
