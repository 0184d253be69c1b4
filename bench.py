#!/usr/bin/env python
"""bench.py -- pair evaluations per second of the periodic-box pair loop (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg4|cfg5|...]

One "step" = one complete pass of the hot path over the synthetic catalogue: cell-list build (key, radix sort,
gather, offsets) + pair kernel + fixed-order reduction (+ for N > 1 the NCCL combine), inputs resident in HBM.
value = binned ordered pairs (sum of DD with unit weights) processed by the whole job per second.
e2e    = the same count divided by the wall time of the public API call (MeasureIABox.measure_xi_w with HOST numpy
         inputs: host preparation, H2D copies, the operator, D2H, post-processing and the HDF5 write are all inside).
The default (headline) workload is BASELINE.json configs[1]: 1e6 galaxies, L = 205, r_p in [0.1, 20], 10 x 8 bins,
wgg + wg+ with 27 jackknife regions (2.99e10 pairs per step).  The same JSON line carries, under "secondary", short runs
of configs[2] (multipoles, every N), configs[3] (1e7 galaxies, N = 8 only: the north-star target) and the
constructor-default 8 x 20 bins, configs[4] through the public API (cfg5: cross-correlation with weights and masks, three lines
of sight, combined covariance; separate calls and the batched measure_xi_projections) and the light-cone brute loop (SURVEY.md
8(f)-4: raw pairs per second through MeasureIALightcone, numpy-oracle CPU leg), and under "parity_check" a comparison of the timed kernel against the reference-exact
general kernel on the full workload plus (N = 1) against the CPU oracle on the cpu_baseline sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

_REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _REPO)

WORKLOADS = {
	# name: (N, boxsize, kind, num_jk, n_r, n_2)
	"cfg2": (1_000_000, 205.0, "w", 27, 10, 8),
	"cfg2_default_bins": (1_000_000, 205.0, "w", 27, 8, 20),
	"cfg3": (1_000_000, 205.0, "multipoles", 27, 10, 8),
	"cfg3_default_bins": (1_000_000, 205.0, "multipoles", 27, 8, 20),
	"cfg4_multipoles": (10_000_000, 300.0, "multipoles", 64, 10, 8),
	"cfg4": (10_000_000, 300.0, "w", 64, 10, 8),
	"small": (100_000, 205.0, "w", 27, 10, 8),
}
FLOP_PER_PAIR = {"w": 64.0, "multipoles": 80.0}  # algorithmic FP64 work per binned pair


def sample_clocks(stop, out):
	"""nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md clocks line).  ONE nvidia-smi
	process in loop mode (-lms 200) is read line by line: spawning a process per sample forks this (large) interpreter every
	200 ms, which showed up as a few ms of jitter on rank 0's short multi-GPU steps."""
	q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
		 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
	dev = os.environ.get("LOCAL_RANK", "0")
	try:
		proc = subprocess.Popen(["nvidia-smi", "-i", dev, f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
								stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
	except Exception:  # noqa: BLE001
		return
	try:
		import select
		while not stop.is_set():
			r, _, _ = select.select([proc.stdout], [], [], 0.2)
			if r:
				line = proc.stdout.readline()
				if not line:
					break
				out.append([x.strip() for x in line.strip().split(",")])
	finally:
		proc.terminate()
		try:
			proc.wait(timeout=2)
		except Exception:  # noqa: BLE001
			proc.kill()


def summarise_clocks(samples):
	if not samples:
		return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
	import statistics
	sm = [float(s[0]) for s in samples if s[0].replace(".", "").isdigit()]
	mx = [float(s[1]) for s in samples if s[1].replace(".", "").isdigit()]
	names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
	reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in samples if len(s) > 3 + i)]
	return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
			"samples": len(samples)}


class ClockSampler:
	def __enter__(self):
		self.samples, self.stop = [], threading.Event()
		self.th = threading.Thread(target=sample_clocks, args=(self.stop, self.samples), daemon=True)
		self.th.start()
		time.sleep(0.25)  # the sampler process is up (and its fork behind us) before the timed region starts
		return self

	def __exit__(self, *exc):
		self.stop.set()
		self.th.join(timeout=2)

	def summary(self):
		return summarise_clocks(self.samples)


def host_threads():
	"""Threads the CPU arm may use: the affinity mask, not OpenMP's default (torchrun exports OMP_NUM_THREADS=1)."""
	try:
		return max(1, len(os.sched_getaffinity(0)))
	except AttributeError:
		return max(1, os.cpu_count() or 1)


def cpu_sample_size(pyoracle, L, kind, num_jk, n_r, n_2, cores, target_s, n_max, n_probe=40_000):
	"""Galaxies in the bounded CPU sample: the binned pairs grow as n^2, so one short probe run fixes the n whose
	step takes about `target_s` seconds on this host."""
	from measure_ia_b200.synthetic import uniform_box
	n_probe = min(n_probe, n_max)
	t = time.perf_counter()
	pyoracle.measure(uniform_box(n_probe, L, seed=1), kind, num_jk=num_jk, boxsize=L, num_bins_r=n_r, num_bins_pi=n_2,
					 n_threads=cores)
	dt = max(time.perf_counter() - t, 1e-3)
	n = int(n_probe * (target_s / dt) ** 0.5)
	return max(min(n, n_max, 400_000), min(n_probe, n_max))


def run_reference(args, N, L, kind, num_jk, n_r, n_2):
	"""--impl reference: the reference's CPU algorithm (oracle port; the Python reference itself cannot travel to the
	GPU box) on all host threads, each step a bounded sample of the same workload (about 2.5 s of CPU work per step,
	so that --steps 20 --warmup 5 ends within ~1.5 min).  Under torchrun only rank 0 works; the others exit 0."""
	rank = int(os.environ.get("RANK", "0"))
	if rank != 0:
		return
	sys.path.insert(0, os.path.join(_REPO, "oracle"))
	import pyoracle
	from measure_ia_b200.synthetic import uniform_box
	pyoracle.build()
	cores = host_threads()
	n_cpu = args.cpu_sample or cpu_sample_size(pyoracle, L, kind, num_jk, n_r, n_2, cores, 2.5, N)
	n_cpu = min(N, n_cpu)
	data = uniform_box(n_cpu, L, seed=1)
	times, pairs = [], 0
	for it in range(args.warmup + args.steps):
		t = time.perf_counter()
		res = pyoracle.measure(data, kind, num_jk=num_jk, boxsize=L, num_bins_r=n_r, num_bins_pi=n_2, n_threads=cores)
		dt = time.perf_counter() - t
		if it >= args.warmup:
			times.append(dt)
			pairs = int(res["__meta__/count"].sum())
	total = sum(times)
	value = pairs * len(times) / total
	sample = f"{n_cpu} galaxies of the same generator, {pairs} binned pairs per step, C/OpenMP restatement (oracle/oracle.c)"
	line = {
		"impl": "reference", "metric": "pair_evals_per_sec", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
		"steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
		"scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
		"config": {"workload": args.workload, "n_galaxies": N, "boxsize": L, "statistic": kind, "num_jk": num_jk,
				   "bins": [n_r, n_2], "sample": sample},
		"cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
		"e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
		"gpu_launches": 0,
	}
	print(json.dumps(line))


class Workload:
	"""One synthetic catalogue resident in HBM and the operator call on it."""

	def __init__(self, name, dev, rank, world, kernel="auto"):
		import numpy as np
		import torch
		from measure_ia_b200 import MeasureIABox, ops
		from measure_ia_b200.synthetic import uniform_box
		self.name = name
		self.N, self.L, self.kind, self.num_jk, self.n_r, self.n_2 = WORKLOADS[name]
		self.dev, self.rank, self.world, self.ops, self.torch = dev, rank, world, ops, torch
		self.data = uniform_box(self.N, self.L, seed=1)
		box = MeasureIABox(self.data, None, boxsize=self.L, num_bins_r=self.n_r, num_bins_pi=self.n_2)
		self.geom = "rppi" if self.kind == "w" else "rmu"
		pos, pos_s, axis, e, w, w_s, same = box._prepare(None, "distortion")
		Lsub = round(self.num_jk ** (1 / 3)) if self.num_jk else 0
		jk = box._jackknife_labels(pos, Lsub).astype(np.int32) if self.num_jk else None
		r2_thr, thr2, rp2_cut, self.clean = box._thresholds_for(self.geom, None)
		self.d_pos = torch.from_numpy(pos).to(dev)
		self.d_jk = torch.from_numpy(jk).to(dev) if jk is not None else None
		self.d_axis, self.d_e = torch.from_numpy(axis).to(dev), torch.from_numpy(e).to(dev)
		self.t_r2, self.t_2 = torch.from_numpy(r2_thr), torch.from_numpy(thr2)
		self.r_search, self.rp2_cut = float(box.r_bins[-1]), float(rp2_cut)
		self.kernel = kernel

	def step(self, kernel=None):
		from measure_ia_b200.box import combine_across_ranks
		ops, torch = self.ops, self.torch
		out = torch.ops.measure_ia_b200.paircount(
			self.d_pos, None, self.d_jk, self.d_pos, None, self.d_jk, self.d_axis, self.d_e, self.t_r2, self.t_2,
			ops.GEOM_RPPI if self.geom == "rppi" else ops.GEOM_RMU, 2, True, self.num_jk, self.L, self.r_search, self.rp2_cut,
			ops.KERNEL_NAMES[kernel or self.kernel], self.rank, self.world)
		if self.world > 1:
			out = combine_across_ranks(*out)
		return out

	def timed(self, steps, warmup, flush, barrier):
		"""W warm-up steps, then K timed steps (CUDA events on the launching stream, L2 flushed between steps,
		barrier + synchronize on both sides, max over ranks)."""
		import torch.distributed as dist
		torch, ops, dev, world = self.torch, self.ops, self.dev, self.world
		for _ in range(warmup):
			out = self.step()
		barrier()
		ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
		kernel_ms, build_ms, reduce_ms = [], [], []
		with ClockSampler() as clk:
			barrier()
			wall0 = time.perf_counter()
			for k in range(steps):
				flush.zero_()  # L2 flush between timed iterations (outside the per-step events)
				ev[k][0].record()
				out = self.step()
				ev[k][1].record()
				# phase times measured by the library with CUDA events on the launching stream (mia_b200.h, timings_host)
				build_ms.append(ops.LAST_TIMINGS_MS[0])
				kernel_ms.append(ops.LAST_TIMINGS_MS[1])
				reduce_ms.append(ops.LAST_TIMINGS_MS[2])
			barrier()
		wall = time.perf_counter() - wall0
		step_ms = [a.elapsed_time(b) for a, b in ev]
		t_total = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
		if world > 1:
			dist.all_reduce(t_total, op=dist.ReduceOp.MAX)
		total_ms = float(t_total.item())
		per_rank = torch.tensor([sum(step_ms) / steps, sum(kernel_ms) / steps], dtype=torch.float64, device=dev)
		if world > 1:
			allr = torch.empty(world * 2, dtype=torch.float64, device=dev)
			dist.all_gather_into_tensor(allr, per_rank)
			per_rank = allr
		per_rank = per_rank.view(-1, 2).cpu().tolist()
		dd_count, stats = out[0], out[7]
		pairs = int(dd_count.sum().item())
		res = {
			"out": out, "pairs": pairs, "tested": int(stats[0].item()), "nan_rule": int(stats[2].item()),
			"kernel_used": int(stats[4].item()), "launches": int(stats[7].item()), "total_ms": total_ms, "wall": wall,
			"value": pairs * steps / (total_ms * 1e-3), "ms_per_step": total_ms / steps,
			"per_rank_ms": [{"step": a, "pair_kernel": b} for a, b in per_rank],
			"phases_ms": {"cell_list_build": sum(build_ms) / steps, "pair_kernel": sum(kernel_ms) / steps,
						  "reductions": sum(reduce_ms) / steps},
			"clocks": clk.summary(),
		}
		return res

	def parity_vs_general(self, out, repeat_general=True):
		"""The timed (tiled) result on the SAME full workload against
		(a) the general kernel (one thread per shape galaxy, the reference's operation sequence incl. divisions, square roots
		    and its NaN rule): pair counts and jackknife pair counts must be bit-identical; fp64 sums within 1e-10 relative
		    + 1e-11 of the largest bin + the general kernel's OWN noise (it adds with atomics in no fixed order: two runs of
		    it differ by up to ~6e-11 of the largest S x D bin at 3e10 pairs -- measured and reported here -- and its error grows
		    like 1e-15 x the pairs in a bin, which is the floor used);
		(b) when the timed kernel is the symmetric one, the ORDERED tiled kernel -- an independent pair loop that evaluates
		    every ordered pair on its own, also with fixed-order sums: agreement to ~1e-13 of the largest bin."""
		torch = self.torch
		names = ("dd_count", "dd_w", "spd", "scd", "dd_jk_count", "dd_jk_w", "spd_jk")

		def rel(a, b):  # largest |a - b| over the largest |a|, worst array
			w = 0.0
			for i in (1, 2, 3, 5, 6):
				if a[i].numel():
					w = max(w, float(((a[i] - b[i]).abs().max() / a[i].abs().max().clamp_min(1e-300)).item()))
			return w

		ref = self.step(kernel="general")
		exact = bool(torch.equal(out[0], ref[0]) and torch.equal(out[4], ref[4]))
		# the checker's rounding noise: unordered fp64 atomics into one accumulator per (region, bin) measured ~1e-15 x (pairs in the
		# bin) against the fixed-order kernels (2.4e-6 at 2.4e9 pairs per bin, 6e-5 at 1.1e11; two runs of it differ by up to that)
		floor = 4e-15 * float(ref[0].max().item())
		noise = [0.0] * 8
		if repeat_general:
			ref2 = self.step(kernel="general")
			exact = exact and bool(torch.equal(ref2[0], ref[0]))
			noise = [float((ref[i].double() - ref2[i].double()).abs().max().item()) if ref[i].numel() else 0.0 for i in range(7)]
			noise_rel = rel(ref, ref2)
			del ref2
		worst = 0.0
		for i in (1, 2, 3, 5, 6):
			a, b = ref[i], out[i]
			if a.numel():
				tol = 1e-10 * a.abs() + 1e-11 * a.abs().max() + max(2.0 * noise[i], floor)
				worst = max(worst, float(((a - b).abs() / tol).max().item()))
		res = {"against": "general kernel (reference-exact arithmetic, measure_w_box_jk.py:401-461) on the full workload",
			   "dd_and_dd_jk_bit_exact": exact, "sums_within_1e-10": bool(worst <= 1.0), "worst_err_over_tol": worst,
			   "timed_vs_general_rel_to_largest_bin": rel(ref, out),
			   "general_run_to_run_rel_to_largest_bin": noise_rel if repeat_general else None,
			   "nan_rule_pairs": int(out[7][2].item()), "nan_rule_pairs_general": int(ref[7][2].item()),
			   "compared": list(names)}
		del ref
		if int(out[7][4].item()) == self.ops.KERNEL_TILED_SYM:
			ordered = self.step(kernel="tiled_ordered")
			res["symmetric_vs_ordered_kernel"] = {
				"dd_and_dd_jk_bit_exact": bool(torch.equal(out[0], ordered[0]) and torch.equal(out[4], ordered[4])),
				"sums_rel_to_largest_bin": rel(ordered, out)}
		return res


def run_cfg5(dev, rank, world, barrier, kernel):
	"""BASELINE.json configs[4] through the public API: a shape sample x density sample cross-correlation with weights and
	boolean masks (Np = 2e6, Ns = 5e5, masks keeping 70 % / 60 %), measured along the three lines of
	sight as the datasets LOS_x / LOS_y / LOS_z with 27 jackknife regions, then the full jackknife covariance of the three
	projections (measure_jackknife.py:573-648).  Host numpy inputs; everything (H2D, prep, kernels, D2H, HDF5, covariance
	combination) is inside the timed wall clock.  One warm-up pass, one timed pass."""
	import numpy as np
	from measure_ia_b200 import MeasureIABox
	from measure_ia_b200.synthetic import uniform_box
	n_p, n_s, L = 2_000_000, 500_000, 205.0
	data = uniform_box(n_p, L, seed=505, n_shape=n_s, weights=True)
	rng = np.random.default_rng(506)
	masks = {"Position": rng.random(n_p) < 0.7, "Position_shape_sample": rng.random(n_s) < 0.6}
	masks["Axis_Direction"] = masks["q"] = masks["Position_shape_sample"]
	masks["weight"], masks["weight_shape_sample"] = masks["Position"], masks["Position_shape_sample"]
	tmp = tempfile.mkdtemp(prefix="mia_cfg5_")
	box = MeasureIABox(data, os.path.join(tmp, f"cfg5_{rank}.hdf5"), boxsize=L, num_bins_r=10, num_bins_pi=8)
	box.kernel = kernel
	names = ["LOS_x", "LOS_y", "LOS_z"]
	wall, pairs, kernel_ms = 0.0, 0, 0.0
	for timed in (False, True):
		barrier()
		t0 = time.perf_counter()
		pairs, kernel_ms = 0, 0.0
		for los, name in enumerate(names):
			data["LOS"] = los
			box.measure_xi_w(name, "both", num_jk=27, temp_file_path=tmp + "/", masks=dict(masks))
			pairs += int(box.last_result["count"].sum())
			kernel_ms += box.last_stats["phases_ms"]["pairs"]
		if rank == 0:
			for corr in ("w_g_plus", "w_gg"):
				box.create_full_cov_matrix_projections(corr, names, num_box=27)
		barrier()
		wall = time.perf_counter() - t0
	# the same work through the batched call: one catalogue preparation and one file handle for the three
	# projections, the covariance combination included
	try:
		barrier()
		t0 = time.perf_counter()
		box.measure_xi_projections(names, "both", num_jk=27, temp_file_path=tmp + "/", masks=dict(masks), statistics="w")
		barrier()
		bw = time.perf_counter() - t0
		bp = sum(int(r["count"].sum()) for r in box.last_results.values())
		batched = {"wall_s": bw, "value": bp / bw, "pairs_equal": bp == pairs, "t_catalogue_s": box.last_stats["t_catalogue"],
				   "t_write_s": box.last_stats["t_write"], "api": "measure_xi_projections(names, statistics='w')"}
	except Exception as exc:  # noqa: BLE001
		batched = {"error": f"{type(exc).__name__}: {exc}"}
	return {"value": pairs / wall, "unit": "pairs/s", "wall_s": wall, "pairs": pairs, "pair_kernel_ms_rank0": kernel_ms,
			"batched": batched,
			"config": {"n_position": n_p, "n_shape": n_s, "masks": "70 % / 60 % kept", "weights": "U(0.5, 1.5)", "boxsize": L,
					   "num_jk": 27, "bins": [10, 8], "datasets": names,
					   "steps": "3 x measure_xi_w(host numpy dict) + create_full_cov_matrix_projections(w_g_plus, w_gg)"},
			"kernel": box.last_stats["kernel"]}


def run_lightcone(dev, rank, world, barrier, peak, cpu_leg):
	"""the light-cone brute loop `_measure_xi_rp_pi_lightcone_brute` (measure_w_lightcone.py:45-214) through the
	public class, host numpy inputs: 4e5 position x 4e5 shape galaxies on a 30 x 30 degree patch, 0.1 < z < 0.4, 10 x 8 bins,
	|Pi| < 60.  The reference visits all N_p N_s pairs, so the like-for-like rate is RAW pairs per second (N_p N_s / time); the
	kernel computes a separation only for pairs inside the chi window and the sky pre-filter."""
	import numpy as np
	from measure_ia_b200.lightcone import MeasureIALightcone

	def catalogue(n, seed):
		rng = np.random.default_rng(seed)
		return {"RA": rng.uniform(0.0, 30.0, n), "DEC": rng.uniform(-15.0, 15.0, n), "Redshift": rng.uniform(0.1, 0.4, n),
				"RA_shape_sample": rng.uniform(0.0, 30.0, n), "DEC_shape_sample": rng.uniform(-15.0, 15.0, n),
				"Redshift_shape_sample": rng.uniform(0.1, 0.4, n), "e1": rng.normal(0, 0.2, n), "e2": rng.normal(0, 0.2, n)}

	n = 400_000
	obj = MeasureIALightcone(catalogue(n, 77), None, None, [0.1, 20.0], 10, 8, 60.0)
	wall, kernel_ms = 0.0, 0.0
	for _ in range(2):  # one warm-up, one timed
		barrier()
		t0 = time.perf_counter()
		obj._measure_xi_rp_pi_lightcone_brute("All", return_output=True, print_num=False)
		barrier()
		wall, kernel_ms = time.perf_counter() - t0, obj.last_stats["kernel_ms"]
	st, res = obj.last_stats, obj.last_result
	binned = int(res["count"].sum())
	rec = {"value": binned / wall, "unit": "pairs/s", "wall_s": wall, "pair_kernel_ms_rank0": kernel_ms, "pairs": binned,
		   "separations_computed": st["tested"], "raw_pairs": n * n, "raw_pairs_per_s": n * n / wall,
		   "split_s": {"t_prep": st.get("t_prep"), "t_device": st.get("t_device")},
		   "config": {"n_position": n, "n_shape": n, "sky": "30 x 30 deg, 0.1 < z < 0.4", "bins": [10, 8], "pi_max": 60.0,
					  "api": "MeasureIALightcone._measure_xi_rp_pi_lightcone_brute(host numpy dict, return_output=True)"}}
	if peak:
		rec["roofline"] = {"bound": "fp64-alu", "achieved": binned * 64 / (kernel_ms * 1e-3) / 1e12 if kernel_ms else None,
						   "peak": peak, "unit": "TFLOP/s", "kernel_ms": kernel_ms,
						   "note": "64 flop credited per BINNED pair as for the box; separations that end in a range reject "
								   "(separations_computed - pairs) are uncredited work"}
		if kernel_ms:
			rec["roofline"]["frac"] = rec["roofline"]["achieved"] / peak
	if cpu_leg:  # the numpy oracle (the reference's own vectorised-per-galaxy arithmetic) on a bounded subsample, one core
		sys.path.insert(0, os.path.join(_REPO, "oracle"))
		import pylightcone
		m = 12000
		sub = catalogue(m, 78)
		t0 = time.perf_counter()
		pos, h = pylightcone.sample(sub["RA"], sub["DEC"], sub["Redshift"])
		shp, _ = pylightcone.sample(sub["RA_shape_sample"], sub["DEC_shape_sample"], sub["Redshift_shape_sample"], e1=sub["e1"], e2=sub["e2"])
		want = pylightcone.pair_sums("rppi", pos, shp, 0.1, 20.0, obj.r_bins, obj.pi_bins, 10, 8)
		dt = time.perf_counter() - t0
		small = MeasureIALightcone(sub, None, None, [0.1, 20.0], 10, 8, 60.0)
		small._measure_xi_rp_pi_lightcone_brute("All", return_output=True, print_num=False)
		rec["cpu_baseline"] = {"raw_pairs_per_s": m * m / dt, "cores": 1, "kind": "port",
							   "sample": f"{m} x {m} galaxies of the same generator in {dt:.1f} s (oracle/pylightcone.py, numpy)"}
		rec["parity_check"] = {"against": "oracle/pylightcone.py on the CPU sample",
							   "dd_bit_exact": bool(np.array_equal(small.last_result["count"], want["count"])),
							   "pairs": int(want["count"].sum())}
	return rec


def fp64_peak(torch, dev):
	"""Dependent-free DFMA rate of this GPU, measured live (MEASURED_PEAKS.json carries no FP64 figure), with the SM
	clock sampled while the probe runs."""
	import ctypes
	from measure_ia_b200.build import PEAKS_PATH
	peaks = ctypes.CDLL(PEAKS_PATH)
	peaks.mia_peak_fp64_tflops.restype = ctypes.c_double
	peaks.mia_peak_fp64_tflops.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
	blocks = 148 * 8
	scratch = torch.empty(blocks * 256, dtype=torch.float64, device=dev)
	torch.cuda.synchronize(dev)
	with ClockSampler() as clk:
		best = 0.0
		t0 = time.perf_counter()
		while time.perf_counter() - t0 < 1.0:  # ~1 s of back-to-back launches so that nvidia-smi sees the load
			best = max(best, float(peaks.mia_peak_fp64_tflops(5, 4096, scratch.data_ptr(), blocks)))
	return best, clk.summary()


def roofline(res, kind, world, peak, peak_clocks, N, workload):
	t_kernel = res["phases_ms"]["pair_kernel"] * 1e-3
	achieved = res["pairs"] / world * FLOP_PER_PAIR[kind] / t_kernel / 1e12 if t_kernel > 0 else None
	hbm_bytes = 96.0 * N
	return {
		"bound": "fp64-alu", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
		"frac": achieved / peak if peak and achieved else None, "traffic": _ncu_traffic(workload),
		"kernel_ms": t_kernel * 1e3, "peak_probe_clocks": peak_clocks,
		"note": ("pair loop is FP64-issue bound (SURVEY.md 8(d)): achieved = binned pairs x %d flop / pair-kernel time "
				 "(CUDA events around the kernel on its stream, rank 0's share at N > 1); peak = dependent-free DFMA rate "
				 "measured live by libmia_peaks.so (MEASURED_PEAKS.json carries no FP64 figure; nominal 148 SM x 64 x 2 x "
				 "1.965 GHz = 37.2); algorithmic HBM bytes per step = %.3g (%.2e B/pair), HBM peak %s GB/s "
				 "(MEASURED_PEAKS.json) is not the limiter" % (
					 int(FLOP_PER_PAIR[kind]), hbm_bytes, hbm_bytes / max(res["pairs"], 1), _hbm_peak())),
	}


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument("--gpus", type=int, default=1)
	ap.add_argument("--steps", type=int, default=5)
	ap.add_argument("--warmup", type=int, default=3)
	ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
	ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
	ap.add_argument("--kernel", default=os.environ.get("MIA_KERNEL", "auto"))
	ap.add_argument("--cpu-sample", type=int, default=0,
					help="galaxies in the bounded CPU-baseline sample (0 = sized by a probe run to ~2.5 s of CPU work)")
	ap.add_argument("--no-cpu-baseline", action="store_true")
	ap.add_argument("--no-e2e", action="store_true")
	ap.add_argument("--no-secondary", action="store_true")
	ap.add_argument("--no-parity", action="store_true")
	args = ap.parse_args()
	N, L, kind, num_jk, n_r, n_2 = WORKLOADS[args.workload]

	if args.impl == "reference":
		return run_reference(args, N, L, kind, num_jk, n_r, n_2)

	import numpy as np
	import torch
	import torch.distributed as dist

	from measure_ia_b200 import MeasureIABox, ops
	from measure_ia_b200.synthetic import uniform_box

	rank = int(os.environ.get("RANK", "0"))
	world = int(os.environ.get("WORLD_SIZE", "1"))
	local = int(os.environ.get("LOCAL_RANK", "0"))
	if not torch.cuda.is_available():
		raise SystemExit("bench.py needs a CUDA device (the pair-count operator has no CPU fallback)")
	torch.cuda.set_device(local)
	dev = torch.device("cuda", local)
	if world > 1:
		dist.init_process_group("nccl", device_id=dev)
	ops.load_library()

	def barrier():
		if world > 1:
			dist.barrier()
		torch.cuda.synchronize(dev)

	flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

	# ---- the headline workload: synthetic catalogue resident in HBM -------------------------------------------------
	W = Workload(args.workload, dev, rank, world, args.kernel)
	res = W.timed(args.steps, args.warmup, flush, barrier)
	pairs = res["pairs"]

	line = {
		"metric": "pair_evals_per_sec", "value": res["value"], "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
		"warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong",
		"vs_baseline": None, "dtype": "f64", "data": "synthetic",
		"config": {"workload": args.workload, "n_galaxies": N, "boxsize": L, "statistic": kind, "num_jk": num_jk,
				   "bins": [n_r, n_2], "pairs_per_step": pairs, "candidates_tested_per_step": res["tested"],
				   "nan_rule_pairs_per_step": res["nan_rule"],
				   "kernel": ops.KERNEL_REPORTED.get(res["kernel_used"], str(res["kernel_used"])),
				   "parallelism": f"shape-sample shards x{world}" if world > 1 else "single GPU",
				   "l2": "flushed between timed steps (256 MiB memset outside the per-step CUDA events)",
				   "wall_s_timed_region": res["wall"], "thresholds_clean": bool(W.clean)},
		"gpu_launches": res["launches"] * args.steps,
		"per_rank_ms": res["per_rank_ms"],
		"phases_ms": res["phases_ms"],
	}
	if rank == 0:
		line["clocks"] = res["clocks"]

	# ---- parity of the timed kernel on the full workload ---------------------------------------------------------------
	if not args.no_parity:
		pc = W.parity_vs_general(res["out"])
		line["parity_check"] = pc

	# ---- roofline of the dominant kernel (FP64 ALU issue; neither HBM nor tensor cores bound this path) ----------------
	peak, peak_clocks = (None, None)
	if rank == 0:
		peak, peak_clocks = fp64_peak(torch, dev)
		line["roofline"] = roofline(res, kind, world, peak, peak_clocks, N, args.workload)

	# ---- end to end through the public API with host buffers ---------------------------------------------------------------
	if not args.no_e2e:
		tmp = tempfile.mkdtemp(prefix="mia_bench_")
		box2 = MeasureIABox(uniform_box(N, L, seed=1), os.path.join(tmp, f"bench_{rank}.hdf5"), boxsize=L,
							num_bins_r=n_r, num_bins_pi=n_2)
		box2.kernel = args.kernel
		run = box2.measure_xi_w if kind == "w" else box2.measure_xi_multipoles
		# a temporary path selects the reference's tree variants (measure_IA.py:102-131), the call a user normally makes
		run("warm", "both", num_jk=num_jk, temp_file_path=tmp + "/")
		barrier()
		t0 = time.perf_counter()
		e2e_steps = max(1, min(args.steps, 3))
		for _ in range(e2e_steps):
			run("All", "both", num_jk=num_jk, temp_file_path=tmp + "/")
		barrier()
		e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
		if world > 1:
			dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
		nb = n_r * n_2
		h2d = box2.last_stats.get("h2d_bytes", N * (24 + 16 + 8 + 8))
		d2h = nb * 8 * 4 + num_jk * nb * 8 * 3 + 64
		line["e2e"] = {"value": pairs * e2e_steps / float(e2e_t.item()), "unit": "pairs/s", "h2d_bytes_per_step": h2d,
					   "d2h_bytes_per_step": d2h, "wall_s_per_call": float(e2e_t.item()) / e2e_steps,
					   "api": "MeasureIABox.measure_xi_w(host numpy dict) incl. host prep, H2D, operator, D2H, HDF5 write",
					   "split_s": {k: box2.last_stats[k] for k in ("t_prep", "t_device", "t_write") if k in box2.last_stats}}
		del box2

	# ---- CPU baseline (oracle port on the host cores, bounded sample) + oracle parity, rank 0 at N = 1 only ---------------
	if rank == 0 and world == 1 and not args.no_cpu_baseline:
		sys.path.insert(0, os.path.join(_REPO, "oracle"))
		import pyoracle
		cores = host_threads()
		n_cpu = args.cpu_sample or cpu_sample_size(pyoracle, L, kind, num_jk, n_r, n_2, cores, 8.0, N)
		sub = uniform_box(min(N, n_cpu), L, seed=1)
		t0 = time.perf_counter()
		want = pyoracle.measure(sub, kind, num_jk=num_jk, boxsize=L, num_bins_r=n_r, num_bins_pi=n_2, n_threads=cores)
		dt = time.perf_counter() - t0
		cp = int(want["__meta__/count"].sum())
		line["cpu_baseline"] = {"value": cp / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
								"sample": f"{len(sub['Position'])} galaxies of the same generator, {cp} binned pairs in {dt:.1f} s "
										  f"(C/OpenMP restatement oracle/oracle.c; the Python reference measured "
										  f"0.74e6 pairs/s on 1 core, BASELINE.md)"}
		if not args.no_parity:  # the oracle as the checker: the CUDA path on the SAME sample
			boxc = MeasureIABox(sub, None, boxsize=L, num_bins_r=n_r, num_bins_pi=n_2)
			boxc.kernel = args.kernel
			(boxc.measure_xi_w if kind == "w" else boxc.measure_xi_multipoles)("All", "both", num_jk=num_jk, temp_file_path=False)
			line["parity_check"]["oracle_sample"] = {
				"galaxies": len(sub["Position"]), "pairs": cp,
				"dd_bit_exact": bool(np.array_equal(boxc.last_result["count"], want["__meta__/count"]))}

	# ---- secondary workloads: short runs so that BASELINE.json configs[2], [3] appear in the driver-run record ------------
	if not args.no_secondary and args.workload == "cfg2":
		sec = {}
		names = ["cfg3", "cfg2_default_bins"] + (["cfg4"] if world >= 8 else [])
		del W, res
		torch.cuda.empty_cache()
		for name in names:
			try:
				Ws = Workload(name, dev, rank, world, args.kernel)
				k_steps = 2 if name == "cfg4" else max(2, min(args.steps, 5))
				r = Ws.timed(k_steps, 1 if name == "cfg4" else 3, flush, barrier)
				rec = {"value": r["value"], "unit": "pairs/s", "ms_per_step": r["ms_per_step"], "steps": k_steps,
					   "pairs_per_step": r["pairs"], "candidates_tested_per_step": r["tested"],
					   "nan_rule_pairs_per_step": r["nan_rule"], "per_rank_ms": r["per_rank_ms"], "phases_ms": r["phases_ms"],
					   "config": dict(zip(("n_galaxies", "boxsize", "statistic", "num_jk", "n_r", "n_2"), WORKLOADS[name]))}
				if rank == 0:
					rec["clocks"] = r["clocks"]
					rec["roofline"] = roofline(r, Ws.kind, world, peak, peak_clocks, Ws.N, name)
					rec["roofline"].pop("note", None)
				if not args.no_parity and (name != "cfg4" or world >= 8):  # (1e7 galaxies: the general kernel needs 8 GPUs)
					rec["parity_check"] = Ws.parity_vs_general(r["out"])  # (cfg4: 2 x 25 s of the sharded general kernel)
				sec[name] = rec
				del Ws, r
				torch.cuda.empty_cache()
			except Exception as exc:  # noqa: BLE001  (a failed secondary run must not lose the headline line)
				sec[name] = {"error": f"{type(exc).__name__}: {exc}"}
		try:
			sec["cfg5"] = run_cfg5(dev, rank, world, barrier, args.kernel)
		except Exception as exc:  # noqa: BLE001
			sec["cfg5"] = {"error": f"{type(exc).__name__}: {exc}"}
		try:
			sec["lightcone"] = run_lightcone(dev, rank, world, barrier, peak, rank == 0 and world == 1 and not args.no_cpu_baseline)
		except Exception as exc:  # noqa: BLE001
			sec["lightcone"] = {"error": f"{type(exc).__name__}: {exc}"}
		line["secondary"] = sec

	if rank == 0:
		print(json.dumps(line))
	if world > 1:
		dist.destroy_process_group()


def _ncu_traffic(workload):
	"""DRAM bytes per launch of the pair kernel from the committed ncu capture (profiles/*traffic.json), or None."""
	for fn in ("r02_traffic.json", "r01_traffic.json"):
		try:
			return json.load(open(os.path.join(_REPO, "profiles", fn)))[workload]["traffic"]
		except Exception:  # noqa: BLE001
			continue
	return None


def _hbm_peak():
	try:
		return json.load(open(os.path.join(_REPO, "MEASURED_PEAKS.json")))["hbm_gbs"]
	except Exception:  # noqa: BLE001
		return "6650 (fallback)"


if __name__ == "__main__":
	main()
