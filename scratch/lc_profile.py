"""Scratch: the bench's light-cone workload alone (no CPU leg), for ncu captures of k_lightcone."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

dev = torch.device("cuda", 0)
r = bench.run_lightcone(dev, 0, 1, lambda: torch.cuda.synchronize(dev), None, False)
print(json.dumps(r))
