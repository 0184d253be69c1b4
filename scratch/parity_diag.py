"""Which sums differ between the kernels at full cfg2 size, and by how much (general run twice: its own noise)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
from measure_ia_b200 import ops
ops.load_library()
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
W = bench.Workload(name, dev, 0, 1, "auto")
outs = {}
for label, k in (("sym", "tiled"), ("ordered", "tiled_ordered"), ("general1", "general"), ("general2", "general")):
	outs[label] = [t.clone() for t in W.step(kernel=k)]
	print(label, "kernel", int(outs[label][7][4]), "pairs", int(outs[label][0].sum()))
names = ("dd_count", "dd_w", "spd", "scd", "dd_jk_count", "dd_jk_w", "spd_jk")
def cmp(a, b):
	print(f"--- {a} vs {b}")
	for i, n in enumerate(names):
		x, y = outs[a][i].double(), outs[b][i].double()
		d = (x - y).abs()
		j = int(d.argmax())
		print(f"  {n:12s} max|d| {float(d.max()):.3e}  max|a| {float(x.abs().max()):.3e}  rel-to-max {float(d.max() / x.abs().max()):.2e}  at value {float(x.flatten()[j]):.4e}")
cmp("sym", "ordered"); cmp("sym", "general1"); cmp("ordered", "general1"); cmp("general1", "general2")
