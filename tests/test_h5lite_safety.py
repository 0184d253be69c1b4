"""CPU: h5lite does not lose or corrupt what it cannot model: read-only handles cannot mutate the cached tree,
a rewrite keeps the file's permission bits, and a file holding attributes is readable but refuses to be rewritten."""
import os
import struct

import numpy as np
import pytest

from measure_ia_b200 import h5lite


def test_read_only_handle_cannot_write(tmp_path):
	path = str(tmp_path / "a.hdf5")
	with h5lite.File(path, "w") as f:
		f.create_dataset("g/x", data=np.arange(4.0))
	f = h5lite.File(path, "r")
	with pytest.raises(OSError):
		f["g/x"][0] = 7.0
	f.close()
	with h5lite.File(path, "a") as f:  # the cached tree served to the next writable handle is untouched
		assert np.array_equal(f["g/x"][:], np.arange(4.0))
		f["g/x"][0] = 7.0
	with h5lite.File(path, "r") as f:
		assert f["g/x"][0] == 7.0


def test_rewrite_keeps_permission_bits(tmp_path):
	path = str(tmp_path / "b.hdf5")
	with h5lite.File(path, "w") as f:
		f.create_dataset("x", data=np.zeros(3))
	os.chmod(path, 0o640)
	with h5lite.File(path, "a") as f:
		f.create_dataset("y", data=np.ones(3))
	assert (os.stat(path).st_mode & 0o777) == 0o640


def test_file_with_attributes_is_not_rewritten(tmp_path):
	"""Patch an attribute message type (0x000C) over the fill-value message of a dataset header: the file still reads, the
	reader flags it, and flushing a modification raises instead of silently dropping the attribute."""
	path = str(tmp_path / "c.hdf5")
	with h5lite.File(path, "w") as f:
		f.create_dataset("x", data=np.arange(5.0))
	raw = bytearray(open(path, "rb").read())
	reader = h5lite._Reader(bytes(raw))
	seen = {}
	orig = h5lite._Reader._messages

	def spy(self, oh_addr):
		msgs = orig(self, oh_addr)
		seen[oh_addr] = msgs
		return msgs
	h5lite._Reader._messages = spy
	try:
		reader.read_into(h5lite.Group("/", None))
	finally:
		h5lite._Reader._messages = orig
	fill = [body for msgs in seen.values() for (mtype, body, _size, _flags) in msgs if mtype == 0x0005]
	assert fill, "no fill-value message found to patch"
	struct.pack_into("<H", raw, fill[0] - 8, 0x000C)  # message header = type u16, size u16, flags u8, 3 reserved
	open(path, "wb").write(bytes(raw))
	h5lite._RECENT.clear()
	f = h5lite.File(path, "a")
	assert np.array_equal(f["x"][:], np.arange(5.0))
	f.create_dataset("y", data=np.ones(2))
	with pytest.raises(NotImplementedError, match="attributes"):
		f.close()
