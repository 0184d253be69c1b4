"""Simulation constants (box size, h) -- mirrors reference ``src/measureia/Sim_info.py:37-135`` (``SimInfo``).

Only the attributes the periodic-box path and the reference's ``tests/test_sim_input.py`` touch are kept:
``simname, snapshot, snap_group, boxsize, L_0p5, h`` (+ the file-info attributes as ``None``).
"""

# name -> (boxsize in cMpc/h, h)
_TABLE = {
	"TNG100": (75.0, 0.6774),
	"TNG100_2": (75.0, 0.6774),
	"TNG300": (205.0, 0.6774),
	"EAGLE": (100.0 * 0.6777, 0.6777),
	"HorizonAGN": (100.0, 0.704),
}


class SimInfo:
	def __init__(self, sim_name, snapshot, boxsize=None, h=None, file_info=False):
		self.simname = sim_name
		self.N_files = None
		self.fof_folder = None
		self.snap_folder = None
		if snapshot is None:
			self.snapshot = None
			self.snap_group = ""
		else:
			self.snapshot = str(snapshot)
			self.snap_group = f"Snapshot_{self.snapshot}/"
		if type(sim_name) == str:
			self.get_specs()
		else:
			self.boxsize = boxsize
			self.h = h
			self.L_0p5 = None if boxsize is None else boxsize / 2.

	def get_specs(self):
		name = self.simname
		if name in _TABLE:
			self.boxsize, self.h = _TABLE[name]
		elif "FLAMINGO" in name:
			if "L1" in name:
				self.boxsize = 1000.0 * 0.681
			elif "L2p8" in name:
				self.boxsize = 2800.0 * 0.681
			else:
				raise KeyError("Add an L1 or L2p8 suffix to your simname to specify which boxsize is used")
			self.h = 0.681
		elif "COLIBRE" in name:
			if "L4" in name:
				self.boxsize = 400.0 * 0.681
			elif "L2" in name:
				self.boxsize = 200.0 * 0.681
			else:
				raise KeyError("Add an L4 or L2 suffix to your simname to specify which boxsize is used")
			self.h = 0.681
		else:
			raise KeyError(
				"Simulation name not recognised. Choose from [TNG100, TNG100_2, TNG300, EAGLE, HorizonAGN, FLAMINGO_L1, "
				"FLAMINGO_L2p8, COLIBRE_L400, COLIBRE_L200].")
		self.L_0p5 = self.boxsize / 2.0
