"""HDF5 output boundary: same helpers and on-disk layout as the reference (``src/measureia/write_data.py:1-56``,
``read_data.py:86-118``).  Uses ``h5py`` when it is installed, else the bundled pure-Python ``h5lite``."""
import numpy as np

from .sim_info import SimInfo

try:  # pragma: no cover - h5py is absent from the build image
	import h5py as _h5
	BACKEND = "h5py"
except Exception:  # noqa: BLE001
	from . import h5lite as _h5
	BACKEND = "h5lite"


def open_file(name, mode="a"):
	return _h5.File(name, mode)


def write_dataset_hdf5(group, name, data):
	"""Create dataset ``name`` in ``group``, replacing an existing one (write_data.py:14-19)."""
	if name in group:
		del group[name]
	group.create_dataset(name, data=data)


def create_group_hdf5(file, name):
	"""``mkdir -p`` for HDF5 groups (write_data.py:23-56); returns the last group of the path."""
	parts = [p for p in name.split("/") if p != ""]
	node = file
	for p in parts:
		node = node[p] if p in node else node.create_group(p)
	return node


class ReadData(SimInfo):
	"""Result reader used by the reference's tests: ``ReadData(sim, catalogue, snapshot, sub_group, data_path=)``
	then ``.read_cat(name)`` (read_data.py:45-118).  Reading raw simulation snapshots is out of scope."""

	def __init__(self, simulation, catalogue, snapshot, sub_group="", output_file_name=None, data_path="./data/raw/"):
		SimInfo.__init__(self, simulation, snapshot)
		self.catalogue = catalogue
		self.sub_group = sub_group
		self.data_path = data_path + "/"
		self.output_file_name = output_file_name

	def read_cat(self, variable, cut=None):
		f = open_file(f"{self.data_path}{self.catalogue}.hdf5", "r")
		try:
			ds = f[f"{self.snap_group}{self.sub_group}{variable}"]
			data = ds[:] if cut is None else ds[cut[0]:cut[1]]
		finally:
			f.close()
		return np.asarray(data)
