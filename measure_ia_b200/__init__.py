"""measure_ia_b200 -- B200-native pair-counting hot path behind the MeasureIA periodic-box API.

Drop-in for ``measureia.MeasureIABox.measure_xi_w`` / ``measure_xi_multipoles`` (reference
``src/measureia/measure_IA.py:10-262``).  Heavy imports (torch, the CUDA library) happen lazily, on first use of
the pair-count ops, so that the light-weight pieces (``h5lite``, ``SimInfo``, ``synthetic``) import anywhere.
"""

__all__ = ["MeasureIABox", "MeasureIABase", "MeasureJackknife", "SimInfo", "ReadData", "write_dataset_hdf5",
		   "create_group_hdf5"]


def __getattr__(name):
	if name in ("MeasureIABox", "MeasureIABase", "MeasureJackknife"):
		from . import box
		return getattr(box, name)
	if name == "SimInfo":
		from .sim_info import SimInfo
		return SimInfo
	if name in ("ReadData", "write_dataset_hdf5", "create_group_hdf5", "open_file"):
		from . import io
		return getattr(io, name)
	raise AttributeError(name)
