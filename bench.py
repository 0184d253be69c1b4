#!/usr/bin/env python
"""bench.py -- pair evaluations per second of the periodic-box pair loop (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg4|small]

One "step" = one complete pass of the hot path over the synthetic catalogue: cell-list build (key, radix sort,
gather, offsets) + pair kernel + fixed-order reduction (+ for N > 1 the NCCL combine), inputs resident in HBM.
value = binned ordered pairs (sum of DD with unit weights) processed by the whole job per second.
e2e    = the same count divided by the wall time of the public API call (MeasureIABox.measure_xi_w with HOST numpy
         inputs: host preparation, H2D copies, the operator, D2H, post-processing and the HDF5 write are all inside).
The default workload is BASELINE.json configs[1]: 1e6 galaxies, L = 205, r_p in [0.1, 20], 10 x 8 bins, wgg + wg+ with
27 jackknife regions (2.99e10 pairs per step).
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
	"""nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md clocks line)."""
	q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
		 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
	dev = os.environ.get("LOCAL_RANK", "0")
	while not stop.is_set():
		try:
			r = subprocess.run(["nvidia-smi", "-i", dev, f"--query-gpu={q}", "--format=csv,noheader,nounits"],
							   capture_output=True, text=True, timeout=5)
			if r.returncode == 0 and r.stdout.strip():
				out.append([x.strip() for x in r.stdout.strip().splitlines()[0].split(",")])
		except Exception:  # noqa: BLE001
			pass
		stop.wait(0.2)


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


def run_reference(args, N, L, kind, num_jk, n_r, n_2):
	"""--impl reference: the reference's CPU algorithm (oracle port; the Python reference itself cannot travel to the
	GPU box) on all host threads, each step a bounded sample of the same workload."""
	rank = int(os.environ.get("RANK", "0"))
	if rank != 0:
		return
	sys.path.insert(0, os.path.join(_REPO, "oracle"))
	import pyoracle
	from measure_ia_b200.synthetic import uniform_box
	pyoracle.build()
	cores = pyoracle.max_threads()
	n_cpu = min(N, args.cpu_sample)
	data = uniform_box(n_cpu, L, seed=1)
	geom_kind = kind
	times, pairs = [], 0
	for it in range(args.warmup + args.steps):
		t = time.perf_counter()
		res = pyoracle.measure(data, geom_kind, num_jk=num_jk, boxsize=L, num_bins_r=n_r, num_bins_pi=n_2, n_threads=cores)
		dt = time.perf_counter() - t
		if it >= args.warmup:
			times.append(dt)
			pairs = int(res["__meta__/count"].sum())
	total = sum(times)
	value = pairs * len(times) / total
	line = {
		"impl": "reference", "metric": "pair_evals_per_sec", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
		"steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
		"scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
		"config": {"workload": args.workload, "n_galaxies": N, "boxsize": L, "statistic": kind, "num_jk": num_jk,
				   "bins": [n_r, n_2], "sample": f"uniform subsample of {n_cpu} galaxies ({pairs} pairs per step)"},
		"cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port",
						 "sample": f"{n_cpu} galaxies, {pairs} binned pairs per step, OpenMP C restatement (oracle/oracle.c)"},
		"e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
		"gpu_launches": 0,
	}
	print(json.dumps(line))


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument("--gpus", type=int, default=1)
	ap.add_argument("--steps", type=int, default=5)
	ap.add_argument("--warmup", type=int, default=3)
	ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
	ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
	ap.add_argument("--kernel", default=os.environ.get("MIA_KERNEL", "auto"))
	ap.add_argument("--cpu-sample", type=int, default=150_000, help="galaxies in the bounded CPU-baseline sample")
	ap.add_argument("--no-cpu-baseline", action="store_true")
	ap.add_argument("--no-e2e", action="store_true")
	args = ap.parse_args()
	N, L, kind, num_jk, n_r, n_2 = WORKLOADS[args.workload]

	if args.impl == "reference":
		return run_reference(args, N, L, kind, num_jk, n_r, n_2)

	import numpy as np
	import torch
	import torch.distributed as dist

	from measure_ia_b200 import MeasureIABox, ops
	from measure_ia_b200.box import combine_across_ranks
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

	# ---- synthetic catalogue, resident in HBM -----------------------------------------------------------------------
	data = uniform_box(N, L, seed=1)
	box = MeasureIABox(data, None, boxsize=L, num_bins_r=n_r, num_bins_pi=n_2)
	box.kernel = args.kernel
	geom = "rppi" if kind == "w" else "rmu"
	pos, pos_s, axis, e, w, w_s, same = box._prepare(None, "distortion")
	Lsub = round(num_jk ** (1 / 3)) if num_jk else 0
	jk = box._jackknife_labels(pos, Lsub).astype(np.int32) if num_jk else None
	r2_thr, thr2, rp2_cut, clean = box._thresholds_for(geom, None)
	d_pos = torch.from_numpy(pos).to(dev)
	d_jk = torch.from_numpy(jk).to(dev) if jk is not None else None
	d_axis, d_e = torch.from_numpy(axis).to(dev), torch.from_numpy(e).to(dev)
	t_r2, t_2 = torch.from_numpy(r2_thr), torch.from_numpy(thr2)
	flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

	def step():
		out = torch.ops.measure_ia_b200.paircount(
			d_pos, None, d_jk, d_pos, None, d_jk, d_axis, d_e, t_r2, t_2, ops.GEOM_RPPI if geom == "rppi" else ops.GEOM_RMU,
			2, True, num_jk, L, float(box.r_bins[-1]), float(rp2_cut), ops.KERNEL_NAMES[args.kernel], rank, world)
		if world > 1:
			out = combine_across_ranks(*out)
		return out

	def barrier():
		if world > 1:
			dist.barrier()
		torch.cuda.synchronize(dev)

	for _ in range(args.warmup):
		out = step()
	barrier()
	clocks, stop = [], threading.Event()
	th = threading.Thread(target=sample_clocks, args=(stop, clocks), daemon=True)
	th.start()
	ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
	wall0 = time.perf_counter()
	kernel_ms, build_ms, reduce_ms = [], [], []
	for k in range(args.steps):
		flush.zero_()  # L2 flush between timed iterations (outside the per-step events)
		ev[k][0].record()
		out = step()
		ev[k][1].record()
		# phase times measured by the library with CUDA events on the launching stream (include/mia_b200.h, timings_host)
		build_ms.append(ops.LAST_TIMINGS_MS[0])
		kernel_ms.append(ops.LAST_TIMINGS_MS[1])
		reduce_ms.append(ops.LAST_TIMINGS_MS[2])
	barrier()
	wall = time.perf_counter() - wall0
	stop.set()
	th.join(timeout=2)
	step_ms = [a.elapsed_time(b) for a, b in ev]
	t_total = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
	if world > 1:
		dist.all_reduce(t_total, op=dist.ReduceOp.MAX)
	total_ms = float(t_total.item())
	per_rank = torch.tensor([sum(step_ms) / args.steps, sum(kernel_ms) / len(kernel_ms)], dtype=torch.float64, device=dev)
	if world > 1:
		allr = torch.empty(world * 2, dtype=torch.float64, device=dev)
		dist.all_gather_into_tensor(allr, per_rank)
		per_rank = allr
	per_rank = per_rank.view(-1, 2).cpu().tolist()
	dd_count, stats = out[0], out[7]
	pairs = int(dd_count.sum().item())
	tested = int(stats[0].item())
	kernel_used = int(stats[4].item())
	value = pairs * args.steps / (total_ms * 1e-3)

	line = {
		"metric": "pair_evals_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
		"warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
		"vs_baseline": None, "dtype": "f64", "data": "synthetic",
		"config": {"workload": args.workload, "n_galaxies": N, "boxsize": L, "statistic": kind, "num_jk": num_jk,
				   "bins": [n_r, n_2], "pairs_per_step": pairs, "candidates_tested_per_step": tested,
				   "kernel": {1: "general", 2: "tiled"}.get(kernel_used, str(kernel_used)),
				   "parallelism": f"shape-sample shards x{world}" if world > 1 else "single GPU",
				   "l2": "flushed between timed steps (256 MiB memset outside the per-step CUDA events)",
				   "wall_s_timed_region": wall, "thresholds_clean": bool(clean)},
		"gpu_launches": int(stats[7].item()) * args.steps,
		"per_rank_ms": [{"step": a, "pair_kernel": b} for a, b in per_rank],
		"phases_ms": {"cell_list_build": sum(build_ms) / len(build_ms), "pair_kernel": sum(kernel_ms) / len(kernel_ms),
					  "reductions": sum(reduce_ms) / len(reduce_ms)},
	}

	if rank == 0:
		line["clocks"] = summarise_clocks(clocks)

	# ---- roofline of the dominant kernel (FP64 ALU issue; neither HBM nor tensor cores bound this path) ----------------
	if rank == 0:
		import ctypes
		from measure_ia_b200.build import PEAKS_PATH
		peaks = ctypes.CDLL(PEAKS_PATH)
		peaks.mia_peak_fp64_tflops.restype = ctypes.c_double
		peaks.mia_peak_fp64_tflops.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
		blocks = 148 * 8
		scratch = torch.empty(blocks * 256, dtype=torch.float64, device=dev)
		torch.cuda.synchronize(dev)
		fp64_peak = float(peaks.mia_peak_fp64_tflops(5, 4096, scratch.data_ptr(), blocks))
		# duration of the dominant (pair) kernel alone, CUDA events on its stream, average over the timed steps
		t_kernel = sum(kernel_ms) / len(kernel_ms) * 1e-3
		achieved = pairs / world * FLOP_PER_PAIR[kind] / t_kernel / 1e12
		hbm_bytes = 96.0 * N
		line["roofline"] = {
			"bound": "fp64-alu", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
			"frac": achieved / fp64_peak if fp64_peak > 0 else None, "traffic": _ncu_traffic(args.workload),
			"kernel_ms": t_kernel * 1e3,
			"note": ("pair loop is FP64-issue bound (SURVEY.md 8(d)): achieved = binned pairs x %d flop / pair-kernel time "
					 "(CUDA events around the kernel on its stream); peak = dependent-free DFMA rate measured live by "
					 "libmia_peaks.so (MEASURED_PEAKS.json carries no FP64 figure); algorithmic HBM bytes per step = %.3g "
					 "(%.2e B/pair), HBM peak %s GB/s (MEASURED_PEAKS.json) is not the limiter" % (
						 int(FLOP_PER_PAIR[kind]), hbm_bytes, hbm_bytes / max(pairs, 1), _hbm_peak())),
		}

	# ---- end to end through the public API with host buffers ---------------------------------------------------------------
	if not args.no_e2e:
		tmp = tempfile.mkdtemp(prefix="mia_bench_")
		box2 = MeasureIABox(uniform_box(N, L, seed=1), os.path.join(tmp, f"bench_{rank}.hdf5"), boxsize=L,
							num_bins_r=n_r, num_bins_pi=n_2)
		box2.kernel = args.kernel
		run = box2.measure_xi_w if kind == "w" else box2.measure_xi_multipoles
		run("warm", "both", num_jk=num_jk, temp_file_path=False)
		barrier()
		t0 = time.perf_counter()
		e2e_steps = max(1, min(args.steps, 3))
		for _ in range(e2e_steps):
			run("All", "both", num_jk=num_jk, temp_file_path=False)
		barrier()
		e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
		if world > 1:
			dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
		nb = n_r * n_2
		h2d = N * (24 + 16 + 8 + 8)  # Position, Axis_Direction, q, weight (jackknife labels are computed on the device)
		d2h = nb * 8 * 4 + num_jk * nb * 8 * 3 + 64
		line["e2e"] = {"value": pairs * e2e_steps / float(e2e_t.item()), "unit": "pairs/s", "h2d_bytes_per_step": h2d,
					   "d2h_bytes_per_step": d2h, "wall_s_per_call": float(e2e_t.item()) / e2e_steps,
					   "api": "MeasureIABox.measure_xi_w(host numpy dict) incl. host prep, H2D, operator, D2H, HDF5 write",
					   "split_s": {k: box2.last_stats[k] for k in ("t_prep", "t_device", "t_write")}}

	# ---- CPU baseline (oracle port on the host cores, bounded sample), rank 0 at N = 1 only ----------------------------------
	if rank == 0 and world == 1 and not args.no_cpu_baseline:
		sys.path.insert(0, os.path.join(_REPO, "oracle"))
		import pyoracle
		cores = pyoracle.max_threads()
		n_cpu = min(N, args.cpu_sample)
		sub = uniform_box(n_cpu, L, seed=1)
		t0 = time.perf_counter()
		res = pyoracle.measure(sub, kind, num_jk=num_jk, boxsize=L, num_bins_r=n_r, num_bins_pi=n_2, n_threads=cores)
		dt = time.perf_counter() - t0
		cp = int(res["__meta__/count"].sum())
		line["cpu_baseline"] = {"value": cp / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
								"sample": f"{n_cpu} galaxies of the same generator, {cp} binned pairs in {dt:.1f} s "
										  f"(C/OpenMP restatement oracle/oracle.c; the Python reference measured "
										  f"0.74e6 pairs/s on 1 core, BASELINE.md)"}
	if rank == 0:
		print(json.dumps(line))
	if world > 1:
		dist.destroy_process_group()


def _ncu_traffic(workload):
	"""DRAM bytes per launch of the pair kernel from the committed ncu capture (profiles/r01_traffic.json), or None."""
	try:
		return json.load(open(os.path.join(_REPO, "profiles", "r01_traffic.json")))[workload]["traffic"]
	except Exception:  # noqa: BLE001
		return None


def _hbm_peak():
	try:
		return json.load(open(os.path.join(_REPO, "MEASURED_PEAKS.json")))["hbm_gbs"]
	except Exception:  # noqa: BLE001
		return "6650 (fallback)"


if __name__ == "__main__":
	main()
