"""CPU, world_size 2, gloo: the exchange step of the sharded path (measure_ia_b200.box.combine_across_ranks).

Reference analogue: the parent-side sum of worker results, measure_w_box_jk.py:775-780.  Integer pair counts are
all-reduced (exact), fp64 sums are all-gathered and added in rank order, so every rank ends with bit-identical arrays."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
	with socket.socket() as s:
		s.bind(("127.0.0.1", 0))
		return s.getsockname()[1]


def _partials(rank, n_r=5, n_2=4, num_jk=8, tasks_skew=0):
	g = torch.Generator().manual_seed(100 + rank)
	dd_count = torch.randint(0, 1000, (n_r, n_2), generator=g, dtype=torch.int64)
	jk_count = torch.randint(0, 100, (num_jk, n_r, n_2), generator=g, dtype=torch.int64)
	f = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64) * 1e3  # noqa: E731
	stats = torch.tensor([10 + rank, 20 + rank, rank, 0, 2, 77, 5 + tasks_skew * rank, 9], dtype=torch.int64)
	return dd_count, f(n_r, n_2), f(n_r, n_2), f(n_r, n_2), jk_count, f(num_jk, n_r, n_2), f(num_jk, n_r, n_2), stats


def _worker(rank, world, port, out_dir, tasks_skew=0, num_jk=8):
	os.environ["MASTER_ADDR"] = "127.0.0.1"
	os.environ["MASTER_PORT"] = str(port)
	dist.init_process_group("gloo", rank=rank, world_size=world)
	try:
		from measure_ia_b200.box import combine_across_ranks
		res = combine_across_ranks(*_partials(rank, tasks_skew=tasks_skew, num_jk=num_jk))
		torch.save([t.clone() for t in res], os.path.join(out_dir, f"rank{rank}.pt"))
	finally:
		dist.destroy_process_group()


def test_combine_across_two_ranks(tmp_path):
	world = 2
	mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
	got = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
	parts = [_partials(r) for r in range(world)]
	names = ["dd_count", "dd_w", "spd", "scd", "jk_count", "jk_w", "spd_jk", "stats"]
	for i, name in enumerate(names[:-1]):
		want = parts[0][i] + parts[1][i]  # rank order: ((rank0) + rank1)
		for r in range(world):
			assert torch.equal(got[r][i], want), f"{name} on rank {r}"
	# statistics: additive entries are summed, the kernel id / cell count are kept
	for r in range(world):
		s = got[r][7]
		assert s[0] == 21 and s[1] == 41 and s[6] == 5 and s[4] == 2 and s[5] == 77


def test_combine_without_jackknife_rows(tmp_path):
	"""The light-cone operator without patches (and the box without jackknife) hands over EMPTY [0, n_r, n_2] jackknife arrays:
	the packed exchange must carry them through (position-sample shards of MeasureIALightcone._pair_sums)."""
	world = 2
	mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), 0, 0), nprocs=world, join=True)
	got = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
	parts = [_partials(r, num_jk=0) for r in range(world)]
	for r in range(world):
		assert len(got[r]) == 8 and got[r][4].shape == (0, 5, 4) and got[r][5].shape == (0, 5, 4)
		for i in range(4):
			assert torch.equal(got[r][i], parts[0][i] + parts[1][i])


def test_ranks_with_different_task_tables_fail_loudly(tmp_path):
	"""Ranks that built different task tables (different SM counts / MIA_* variables) would drop or double-count tasks."""
	with pytest.raises(Exception, match="ranks disagree"):
		mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), 1), nprocs=2, join=True)


def test_shard_slices_cover_the_sample():
	"""The general kernel shards the cell-sorted shape sample by index: [n*i/w, n*(i+1)/w) must tile [0, n)."""
	for n in (0, 1, 7, 1000, 10 ** 6 + 3):
		for w in (1, 2, 3, 8):
			edges = [n * i // w for i in range(w + 1)]
			assert edges[0] == 0 and edges[-1] == n and all(b >= a for a, b in zip(edges, edges[1:]))
