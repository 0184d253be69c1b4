"""Summarise an ncu report: key metrics, stall reasons, and executed instructions / samples by SASS address bucket."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
for vals in rows[2:]:
	d = dict(zip(hdr, vals))
	print("kernel:", d.get("Kernel Name"))
	keys = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
			'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
			'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.per_cycle_active',
			'smsp__average_warp_latency_per_inst_issued.ratio', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
			'launch__occupancy_limit_registers', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
			'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
			'launch__shared_mem_per_block_dynamic', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
			'smsp__inst_executed_op_shared_atom.sum', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
			'smsp__inst_executed_op_global_ld.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']
	for k in keys:
		if k in d:
			print(" ", k, d[k])
	for h in hdr:
		if 'average_warps_issue_stalled' in h and h.endswith('.ratio') and 'not_issued' not in h:
			try:
				v = float(d[h])
				if v > 0.05:
					print("  stall", h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), round(v, 3))
			except Exception:
				pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
iA = hdr.index('Address'); iS = hdr.index('Source'); iE = hdr.index('Instructions Executed'); iN = hdr.index('# Samples')
base = int(data[0][iA], 16)
tot = sum(int(r[iE]) for r in data); tots = sum(int(r[iN]) for r in data)
print('total instr', tot, 'samples', tots)
b = collections.OrderedDict()
B = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0x800
for r in data:
	off = int(r[iA], 16) - base
	e = b.setdefault(off // B, [0, 0]); e[0] += int(r[iE]); e[1] += int(r[iN])
for k, (e, s) in b.items():
	if e / tot > 0.004 or s / tots > 0.004:
		print(hex(k * B), 'instr %.1f%%' % (100 * e / tot), 'samples %.1f%%' % (100 * s / tots))
with open(rep + ".sass.txt", "w") as f:
	for r in data:
		f.write(f"{hex(int(r[iA],16)-base)} {r[iE]:>12} {r[iN]:>7}  {r[iS]}\n")
