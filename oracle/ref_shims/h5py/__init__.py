"""h5py stand-in for running the unmodified reference: forwards to measure_ia_b200.h5lite (real HDF5 files)."""
from measure_ia_b200.h5lite import File, Group, Dataset  # noqa: F401
