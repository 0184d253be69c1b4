"""Minimal pure-Python HDF5 reader/writer with an h5py-like surface.

Why this exists: the reference keeps every result in one HDF5 file (layout in DESIGN.md §boundary, produced by
reference ``src/measureia/write_data.py:1-56`` through ``h5py``), and "the HDF5 output layout stays unchanged" is
part of the drop-in contract -- but neither ``h5py`` nor ``libhdf5`` exists in this image or on the GPU box.
``open_file`` in ``measure_ia_b200.io`` uses ``h5py`` when it is importable and this module otherwise.

Subset implemented (exactly what h5py's default "earliest" file format uses for the reference's outputs, and what
the two golden files under the reference's ``tests/data/processed/TNG300`` contain):

* superblock version 0, 8-byte offsets/lengths
* groups: version-1 object header + symbol-table message, v1 B-tree (node type 0), SNOD leaves, local heap
* datasets: version-1 object header with dataspace v1/v2, datatype v1 (little-endian IEEE floats and integers),
  fill-value message, data layout v3 (contiguous or compact) ; layout v1/v2 contiguous are read as well
* object-header continuation blocks are followed when reading

Chunked / compressed datasets, attributes, links other than hard links and big-endian types are not supported and
raise ``NotImplementedError`` when encountered.  Attributes are skipped on reading; a file that holds any can be read but
not rewritten (every flush re-serialises the whole file, which would drop them): ``flush`` raises instead.

The file is parsed fully on open and re-serialised on ``close()`` when it was modified (the reference's files are a
few hundred small float64 arrays, well under a few MB), which gives h5py's ``'a'`` semantics: existing content is
kept, ``del group[name]`` followed by ``create_dataset`` overwrites.
"""
from __future__ import annotations

import os
import struct
from collections import OrderedDict

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF
_LEAF_K = 32  # symbol-table leaf K written into our superblocks (2K = 64 entries per SNOD)
_INT_K = 16  # B-tree internal K (2K = 32 children per node)


# ----------------------------------------------------------------------------------------------------------------
# in-memory object model
# ----------------------------------------------------------------------------------------------------------------
def _split(path):
	if isinstance(path, bytes):
		path = path.decode()
	return [p for p in str(path).split("/") if p != ""]


class Dataset:
	"""A dataset held fully in memory (numpy array)."""

	def __init__(self, name, data, parent=None):
		self._data = np.array(data)  # copy: later edits of the caller's array must not leak into the file
		self.name = name
		self.parent = parent

	@property
	def shape(self):
		return self._data.shape

	@property
	def dtype(self):
		return self._data.dtype

	@property
	def size(self):
		return self._data.size

	@property
	def ndim(self):
		return self._data.ndim

	def __len__(self):
		if self._data.ndim == 0:
			raise TypeError("Attempt to take len() of scalar dataset")
		return self._data.shape[0]

	def __getitem__(self, sel):
		out = self._data[sel]
		return np.array(out) if isinstance(out, np.ndarray) else out

	def __setitem__(self, sel, value):
		node = self
		while node.parent is not None:
			node = node.parent
		if getattr(node, "mode", "a") == "r":  # (a cached tree may be shared with a later writable handle)
			raise OSError("Unable to write to dataset (file opened read-only)")
		self._data[sel] = value
		node._dirty = True

	def __array__(self, dtype=None, copy=None):
		arr = self._data if dtype is None else self._data.astype(dtype)
		return np.array(arr) if copy else arr

	def __iter__(self):
		return iter(self._data)

	def __repr__(self):
		return f'<h5lite dataset "{self.name}": shape {self.shape}, type "{self.dtype.str}">'


class Group:
	def __init__(self, name="/", parent=None):
		self.name = name
		self.parent = parent
		self._children = OrderedDict()

	# -- navigation -------------------------------------------------------------------------------------------
	def _root(self):
		node = self
		while node.parent is not None:
			node = node.parent
		return node

	def _walk(self, path, create=False):
		start = self._root() if (isinstance(path, str) and path.startswith("/")) else self
		node = start
		for part in _split(path):
			if not isinstance(node, Group):
				raise KeyError(f"Unable to open object (component not a group: {part!r})")
			if part not in node._children:
				if not create:
					raise KeyError(f"Unable to open object (object '{part}' doesn't exist)")
				child = Group((node.name.rstrip("/") + "/" + part), node)
				node._children[part] = child
				self._root()._dirty = True
			node = node._children[part]
		return node

	def __getitem__(self, path):
		return self._walk(path)

	def __contains__(self, path):
		try:
			self._walk(path)
			return True
		except KeyError:
			return False

	def get(self, path, default=None):
		try:
			return self._walk(path)
		except KeyError:
			return default

	def __delitem__(self, path):
		parts = _split(path)
		if not parts:
			raise KeyError("cannot delete the group itself")
		parent = self._walk("/".join(parts[:-1])) if len(parts) > 1 else (
			self._root() if str(path).startswith("/") else self)
		if parts[-1] not in parent._children:
			raise KeyError(f"Couldn't delete link (name doesn't exist: {parts[-1]!r})")
		del parent._children[parts[-1]]
		self._root()._dirty = True

	def keys(self):
		return list(self._children.keys())

	def values(self):
		return list(self._children.values())

	def items(self):
		return list(self._children.items())

	def __iter__(self):
		return iter(self.keys())

	def __len__(self):
		return len(self._children)

	# -- creation ---------------------------------------------------------------------------------------------
	def create_group(self, path):
		parts = _split(path)
		if not parts:
			raise ValueError("Unable to create group (name already exists)")
		parent = self._walk("/".join(parts[:-1]), create=True) if len(parts) > 1 else (
			self._root() if str(path).startswith("/") else self)
		if parts[-1] in parent._children:
			raise ValueError("Unable to create group (name already exists)")
		g = Group(parent.name.rstrip("/") + "/" + parts[-1], parent)
		parent._children[parts[-1]] = g
		self._root()._dirty = True
		return g

	def require_group(self, path):
		node = self._walk(path, create=True)
		if not isinstance(node, Group):
			raise TypeError("Incompatible object (Dataset) already exists")
		return node

	def create_dataset(self, name, shape=None, dtype=None, data=None):
		parts = _split(name)
		if not parts:
			raise ValueError("dataset needs a name")
		parent = self._walk("/".join(parts[:-1]), create=True) if len(parts) > 1 else (
			self._root() if str(name).startswith("/") else self)
		if parts[-1] in parent._children:
			raise ValueError("Unable to create dataset (name already exists)")
		if data is None:
			arr = np.zeros(shape if shape is not None else (), dtype=dtype or np.float64)
		else:
			arr = np.array(data, dtype=dtype, copy=True)  # own the payload (h5py copies too): later in-place changes of the
			# caller's array must not leak into the tree, which may be reused by the next handle on this file
			if shape is not None:
				arr = arr.reshape(shape)
		if arr.dtype == np.bool_:
			arr = arr.astype(np.int8)  # h5py stores bools as an enum over int8; we keep the payload
		if arr.dtype.kind not in "fiu":
			raise TypeError(f"h5lite can only store integer / float arrays, got {arr.dtype}")
		ds = Dataset(parent.name.rstrip("/") + "/" + parts[-1], arr, parent)
		parent._children[parts[-1]] = ds
		self._root()._dirty = True
		return ds

	def __repr__(self):
		return f'<h5lite group "{self.name}" ({len(self._children)} members)>'


_OPEN_FILES = {}  # realpath -> shared state of every handle currently open on that file in this process
# realpath -> (stat signature, tree) of files this process wrote last: re-opening a file that has not changed on disk since
# (same inode, size and mtime_ns) reuses the tree instead of parsing ~50 us per dataset again.  A measurement appends to the
# same output file call after call (reference usage: one file per simulation), so this halves the cost of the write phase.
_RECENT = {}
_RECENT_MAX = 4


def _stat_sig(path):
	st = os.stat(path)
	return (st.st_ino, st.st_size, st.st_mtime_ns)


class File(Group):
	"""``h5lite.File(name, mode)`` with modes 'r', 'r+', 'a', 'w' (h5py semantics).

	Handles opened on the same path share one in-memory tree (as handles of one process share one file in libhdf5), so a
	handle that is never closed -- the reference leaks one in ``measure_jackknife.py:603`` -- cannot overwrite later
	changes with a stale copy when it is finally collected."""

	def __init__(self, filename, mode="r"):
		super().__init__("/", None)
		self.filename = str(filename)
		self.mode = mode
		self._open = True
		key = os.path.realpath(self.filename)
		exists = os.path.exists(self.filename)
		if mode in ("r", "r+") and not exists and key not in _OPEN_FILES:
			raise FileNotFoundError(f"Unable to open file (unable to open file: name = '{self.filename}')")
		if mode == "w-" and exists:
			raise FileExistsError(self.filename)
		if mode not in ("r", "r+", "a", "w", "w-", "x"):
			raise ValueError(f"invalid mode {mode!r}")
		shared = _OPEN_FILES.get(key)
		if shared is not None and mode in ("r", "r+", "a"):
			self._shared = shared
			self._children = shared["children"]
			shared["open"] += 1
		else:
			self._shared = {"children": self._children, "dirty": False, "open": 1, "key": key}
			_OPEN_FILES[key] = self._shared
			if mode in ("r", "r+", "a") and exists and os.path.getsize(self.filename) > 0:
				recent = _RECENT.get(key)
				if recent is not None and recent[0] == _stat_sig(self.filename):
					self._children = self._shared["children"] = recent[1]
					for child in self._children.values():  # the tree now hangs off THIS handle (dirty flag, absolute paths)
						child.parent = self
				else:
					with open(self.filename, "rb") as fh:
						reader = _Reader(fh.read())
						reader.read_into(self)
						if reader.unsupported:
							self._shared["lossy"] = ", ".join(sorted(reader.unsupported))
				self._shared["dirty"] = False
			else:
				self._shared["dirty"] = True  # a new (possibly empty) file must still be written

	@property
	def _dirty(self):
		return self._shared["dirty"]

	@_dirty.setter
	def _dirty(self, value):
		self._shared["dirty"] = value

	def flush(self):
		if self._shared["dirty"] and (self.mode != "r" or self._shared["open"] > 1):
			if self._shared.get("lossy"):
				# every flush re-serialises the whole file: content h5lite does not model would silently disappear
				raise NotImplementedError(f"{self.filename} holds {self._shared['lossy']}, which h5lite cannot preserve when "
										  "rewriting the file; write to a new file (or install h5py)")
			blob = _Writer().serialise(self)
			tmp = self.filename + ".h5lite.tmp"
			with open(tmp, "wb") as fh:
				fh.write(blob)
			if os.path.exists(self.filename):  # keep the permission bits of the file being replaced
				try:
					os.chmod(tmp, os.stat(self.filename).st_mode & 0o7777)
				except OSError:
					pass
			os.replace(tmp, self.filename)
			self._shared["dirty"] = False
			_RECENT.pop(self._shared["key"], None)
			while len(_RECENT) >= _RECENT_MAX:
				_RECENT.pop(next(iter(_RECENT)))
			_RECENT[self._shared["key"]] = (_stat_sig(self.filename), self._shared["children"])

	def close(self):
		if self._open:
			self.flush()
			self._open = False
			self._shared["open"] -= 1
			if self._shared["open"] <= 0 and _OPEN_FILES.get(self._shared["key"]) is self._shared:
				del _OPEN_FILES[self._shared["key"]]

	def __enter__(self):
		return self

	def __exit__(self, *exc):
		self.close()
		return False

	def __del__(self):
		try:
			self.close()
		except Exception:
			pass


# ----------------------------------------------------------------------------------------------------------------
# reader
# ----------------------------------------------------------------------------------------------------------------
class _Reader:
	def __init__(self, buf):
		self.b = buf
		self.unsupported = set()  # content that is skipped on reading and therefore lost when the file is rewritten
		if buf[:8] != _SIG:
			raise OSError("not an HDF5 file (bad signature)")
		ver = buf[8]
		if ver not in (0, 1):
			raise NotImplementedError(f"h5lite reads superblock version 0/1 only (found {ver})")
		so, sl = buf[13], buf[14]
		if (so, sl) != (8, 8):
			raise NotImplementedError("h5lite needs 8-byte offsets and lengths")
		p = 24 if ver == 0 else 28
		self.base, _fs, self.eof, _drv = struct.unpack_from("<QQQQ", buf, p)
		p += 32
		# root group symbol table entry
		_lno, self.root_oh, cache, _r = struct.unpack_from("<QQII", buf, p)
		self.root_scratch = struct.unpack_from("<QQ", buf, p + 24)

	def read_into(self, root):
		self._read_group(self.root_oh, root)

	# -- object headers ---------------------------------------------------------------------------------------
	def _messages(self, addr):
		b = self.b
		ver = b[addr]
		if ver != 1:
			raise NotImplementedError(f"object header version {ver} (only v1 supported)")
		nmsg, _ref, hsize = struct.unpack_from("<HII", b, addr + 2)
		blocks = [(addr + 16, hsize)]
		msgs = []
		while blocks and len(msgs) < nmsg:
			p, left = blocks.pop(0)
			end = p + left
			while p + 8 <= end and len(msgs) < nmsg:
				mtype, msize, mflags = struct.unpack_from("<HHB", b, p)
				body = p + 8
				if mtype in (0x000C, 0x0015):
					self.unsupported.add("attributes")
				if mtype == 0x0010:  # continuation
					caddr, clen = struct.unpack_from("<QQ", b, body)
					blocks.append((caddr, clen))
				msgs.append((mtype, body, msize, mflags))
				p = body + msize
		return msgs

	def _read_group(self, oh_addr, group):
		stab = None
		for mtype, body, msize, _f in self._messages(oh_addr):
			if mtype == 0x0011:
				stab = struct.unpack_from("<QQ", self.b, body)
			elif mtype in (0x0002, 0x0006):
				raise NotImplementedError("new-style (link message) groups are not supported by h5lite")
		if stab is None:
			raise OSError("group without symbol table message")
		btree, heap = stab
		hb = self.b
		if hb[heap:heap + 4] != b"HEAP":
			raise OSError("bad local heap signature")
		_dsize, _free, hdata = struct.unpack_from("<QQQ", hb, heap + 8)
		for name_off, oh in self._iter_btree(btree):
			end = hb.index(b"\x00", hdata + name_off)
			name = hb[hdata + name_off:end].decode()
			self._read_object(oh, name, group)

	def _iter_btree(self, addr):
		b = self.b
		if addr == _UNDEF:
			return
		if b[addr:addr + 4] != b"TREE":
			raise OSError("bad B-tree signature")
		ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
		if ntype != 0:
			raise OSError("expected a group B-tree node")
		p = addr + 24
		for i in range(used):
			child = struct.unpack_from("<Q", b, p + 8)[0]
			p += 16
			if level > 0:
				yield from self._iter_btree(child)
			else:
				if b[child:child + 4] != b"SNOD":
					raise OSError("bad symbol table node signature")
				nsym = struct.unpack_from("<H", b, child + 6)[0]
				q = child + 8
				for _ in range(nsym):
					name_off, oh = struct.unpack_from("<QQ", b, q)
					yield name_off, oh
					q += 40

	def _read_object(self, oh_addr, name, parent):
		msgs = self._messages(oh_addr)
		types = {m[0] for m in msgs}
		full = parent.name.rstrip("/") + "/" + name
		if 0x0011 in types:
			g = Group(full, parent)
			parent._children[name] = g
			self._read_group(oh_addr, g)
			return
		shape = dtype = None
		layout = None
		for mtype, body, msize, _f in msgs:
			if mtype == 0x0001:
				shape = self._dataspace(body)
			elif mtype == 0x0003:
				dtype = self._datatype(body)
			elif mtype == 0x0008:
				layout = self._layout(body)
			elif mtype == 0x000B:
				raise NotImplementedError("filtered (compressed) datasets are not supported by h5lite")
		if shape is None or dtype is None or layout is None:
			raise NotImplementedError(f"object '{full}' is neither an old-style group nor a simple dataset")
		count = int(np.prod(shape)) if len(shape) else 1
		kind, a, b_ = layout
		if kind == "contiguous":
			if a == _UNDEF or count == 0:
				arr = np.zeros(shape, dtype=dtype)
			else:
				arr = np.frombuffer(self.b, dtype=dtype, count=count, offset=a).reshape(shape).copy()
		else:  # compact
			arr = np.frombuffer(self.b, dtype=dtype, count=count, offset=a).reshape(shape).copy()
		parent._children[name] = Dataset(full, arr, parent)

	def _dataspace(self, p):
		b = self.b
		ver, rank, flags = b[p], b[p + 1], b[p + 2]
		if ver == 1:
			q = p + 8
		elif ver == 2:
			if b[p + 3] == 2:  # null dataspace
				return (0,)
			q = p + 4
		else:
			raise NotImplementedError(f"dataspace version {ver}")
		return tuple(struct.unpack_from("<" + "Q" * rank, b, q)) if rank else ()

	def _datatype(self, p):
		b = self.b
		cls, ver = b[p] & 0x0F, b[p] >> 4
		bits0 = b[p + 1]
		size = struct.unpack_from("<I", b, p + 4)[0]
		if bits0 & 1:
			raise NotImplementedError("big-endian datatypes are not supported by h5lite")
		if cls == 0:
			signed = bool(bits0 & 0x08)
			return np.dtype(("<i" if signed else "<u") + str(size))
		if cls == 1:
			return np.dtype("<f" + str(size))
		if cls == 8:  # enum (h5py bool): base type follows the 8-byte header
			return self._datatype(p + 8)
		raise NotImplementedError(f"datatype class {cls}")

	def _layout(self, p):
		b = self.b
		ver = b[p]
		if ver == 3:
			cls = b[p + 1]
			if cls == 1:
				addr, size = struct.unpack_from("<QQ", b, p + 2)
				return ("contiguous", addr, size)
			if cls == 0:
				size = struct.unpack_from("<H", b, p + 2)[0]
				return ("compact", p + 4, size)
			raise NotImplementedError("chunked datasets are not supported by h5lite")
		if ver in (1, 2):
			rank, cls = b[p + 1], b[p + 2]
			if cls == 1:
				addr = struct.unpack_from("<Q", b, p + 8)[0]
				return ("contiguous", addr, 0)
			if cls == 0:
				q = p + 8 + 4 * rank
				size = struct.unpack_from("<I", b, q)[0]
				return ("compact", q + 4, size)
			raise NotImplementedError("chunked datasets are not supported by h5lite")
		raise NotImplementedError(f"data layout version {ver}")


# ----------------------------------------------------------------------------------------------------------------
# writer
# ----------------------------------------------------------------------------------------------------------------
def _pad8(n):
	return (n + 7) & ~7


class _Writer:
	def __init__(self):
		self.buf = bytearray()

	def _alloc(self, n, align=8):
		pos = (len(self.buf) + align - 1) // align * align
		self.buf.extend(b"\x00" * (pos + n - len(self.buf)))
		return pos

	def _put(self, pos, data):
		self.buf[pos:pos + len(data)] = data

	def serialise(self, root):
		# superblock v0 is 56 bytes + 40-byte root symbol table entry = 96
		self._alloc(96)
		root_oh, btree, heap = self._write_group(root)
		sb = bytearray()
		sb += _SIG
		sb += bytes([0, 0, 0, 0, 0, 8, 8, 0])
		sb += struct.pack("<HHI", _LEAF_K, _INT_K, 0)
		sb += struct.pack("<QQQQ", 0, _UNDEF, len(self.buf), _UNDEF)
		sb += struct.pack("<QQII", 0, root_oh, 1, 0) + struct.pack("<QQ", btree, heap)
		assert len(sb) == 96
		self._put(0, sb)
		self._put(40, struct.pack("<Q", len(self.buf)))  # end-of-file address
		return bytes(self.buf)

	# -- groups -------------------------------------------------------------------------------------------------
	def _write_group(self, group):
		# children first so their header addresses are known
		entries = []  # (name, oh_addr, cache_type, scratch)
		for name, child in group._children.items():
			if isinstance(child, Group):
				oh, bt, hp = self._write_group(child)
				entries.append((name, oh, 1, struct.pack("<QQ", bt, hp)))
			else:
				oh = self._write_dataset(child)
				entries.append((name, oh, 0, b"\x00" * 16))
		entries.sort(key=lambda e: e[0].encode())

		# local heap: offset 0 holds the empty string (B-tree key 0)
		heap_data = bytearray(b"\x00" * 8)
		name_off = {}
		for name, *_ in entries:
			name_off[name] = len(heap_data)
			raw = name.encode() + b"\x00"
			heap_data += raw + b"\x00" * (_pad8(len(raw)) - len(raw))
		free_off = len(heap_data)
		heap_data += b"\x00" * 16  # one free block at the tail (next = 1 == none, size = 16)
		struct.pack_into("<QQ", heap_data, free_off, 1, 16)
		hdata_addr = self._alloc(len(heap_data))
		self._put(hdata_addr, heap_data)
		heap_addr = self._alloc(32)
		self._put(heap_addr, b"HEAP" + bytes([0, 0, 0, 0]) + struct.pack("<QQQ", len(heap_data), free_off, hdata_addr))

		# symbol table nodes
		per = 2 * _LEAF_K
		leaves = []  # (snod_addr, last_name_offset)
		for i in range(0, len(entries), per):
			chunk = entries[i:i + per]
			addr = self._alloc(8 + per * 40)
			blob = bytearray(b"SNOD" + bytes([1, 0]) + struct.pack("<H", len(chunk)))
			for name, oh, ctype, scratch in chunk:
				blob += struct.pack("<QQII", name_off[name], oh, ctype, 0) + scratch
			self._put(addr, blob)
			leaves.append((addr, name_off[chunk[-1][0]]))

		btree_addr = self._write_btree(leaves, 0)
		# object header: one symbol-table message
		oh_addr = self._alloc(16 + 24)
		oh = bytearray(struct.pack("<BBHII", 1, 0, 1, 1, 24) + b"\x00" * 4)
		oh += struct.pack("<HHB3x", 0x0011, 16, 0) + struct.pack("<QQ", btree_addr, heap_addr)
		self._put(oh_addr, oh)
		return oh_addr, btree_addr, heap_addr

	def _write_btree(self, children, level):
		"""children: list of (address, last-key heap offset). Returns the address of the (sub)tree root node."""
		per = 2 * _INT_K
		node_size = 24 + (2 * per + 1) * 8
		if len(children) <= per:
			addr = self._alloc(node_size)
			blob = bytearray(b"TREE" + struct.pack("<BBH", 0, level, len(children)) + struct.pack("<QQ", _UNDEF, _UNDEF))
			blob += struct.pack("<Q", 0)  # key 0: the empty string
			for caddr, key in children:
				blob += struct.pack("<QQ", caddr, key)
			self._put(addr, blob)
			return addr
		# split into sibling nodes at this level, then index them one level up
		nodes = []
		groups = [children[i:i + per] for i in range(0, len(children), per)]
		addrs = [self._alloc(node_size) for _ in groups]
		prev_key = 0
		for gi, grp in enumerate(groups):
			left = addrs[gi - 1] if gi > 0 else _UNDEF
			right = addrs[gi + 1] if gi + 1 < len(groups) else _UNDEF
			blob = bytearray(b"TREE" + struct.pack("<BBH", 0, level, len(grp)) + struct.pack("<QQ", left, right))
			blob += struct.pack("<Q", prev_key)
			for caddr, key in grp:
				blob += struct.pack("<QQ", caddr, key)
			self._put(addrs[gi], blob)
			prev_key = grp[-1][1]
			nodes.append((addrs[gi], grp[-1][1]))
		return self._write_btree(nodes, level + 1)

	# -- datasets -----------------------------------------------------------------------------------------------
	@staticmethod
	def _datatype_msg(dt):
		dt = np.dtype(dt)
		size = dt.itemsize
		if dt.kind == "f":
			if size == 8:
				props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
				bits = bytes([0x20, 63, 0])
			elif size == 4:
				props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
				bits = bytes([0x20, 31, 0])
			elif size == 2:
				props = struct.pack("<HHBBBBI", 0, 16, 10, 5, 0, 10, 15)
				bits = bytes([0x20, 15, 0])
			else:
				raise TypeError(dt)
			return bytes([0x11]) + bits + struct.pack("<I", size) + props
		if dt.kind in "iu":
			bits = bytes([0x08 if dt.kind == "i" else 0x00, 0, 0])
			return bytes([0x10]) + bits + struct.pack("<I", size) + struct.pack("<HH", 0, size * 8)
		raise TypeError(dt)

	_HEADER_CACHE = {}  # (dtype string, shape) -> (object header with a blank layout address / size, offset of that field)

	@classmethod
	def _dataset_header(cls, dtype, shape):
		key = (dtype.str, shape)
		hit = cls._HEADER_CACHE.get(key)
		if hit is not None:
			return hit
		rank = len(shape)
		# dataspace v1 exactly as h5py writes it: flag bit 0 set, max dims == dims
		dims = b"".join(struct.pack("<Q", d) for d in shape)
		dspace = bytes([1, rank, 1 if rank else 0, 0, 0, 0, 0, 0]) + dims + (dims if rank else b"")
		fill = bytes([2, 2, 2, 1, 0, 0, 0, 0])  # v2, late allocation, fill "if set", defined with size 0 (h5py default)
		layout = bytes([3, 1]) + b"\x00" * 16  # contiguous; address and size are patched per dataset
		msgs = [(0x0001, dspace, 0), (0x0003, cls._datatype_msg(dtype), 1), (0x0005, fill, 1), (0x0008, layout, 0)]
		body = bytearray()
		patch = 0
		for mtype, payload, flags in msgs:
			plen = _pad8(len(payload))
			if mtype == 0x0008:
				patch = 16 + len(body) + 8 + 2  # object header prefix + message header + (version, class)
			body += struct.pack("<HHB3x", mtype, plen, flags) + payload + b"\x00" * (plen - len(payload))
		header = struct.pack("<BBHII", 1, 0, len(msgs), 1, len(body)) + b"\x00" * 4 + bytes(body)
		if len(cls._HEADER_CACHE) > 256:
			cls._HEADER_CACHE.clear()
		cls._HEADER_CACHE[key] = (header, patch)
		return header, patch

	def _write_dataset(self, ds):
		arr = np.ascontiguousarray(ds._data)
		if arr.dtype.byteorder == ">":
			arr = arr.astype(arr.dtype.newbyteorder("<"))
		raw = arr.tobytes()
		data_addr = self._alloc(len(raw)) if len(raw) else _UNDEF
		if len(raw):
			self._put(data_addr, raw)
		header, patch = self._dataset_header(arr.dtype, arr.shape)
		oh_addr = self._alloc(len(header))
		self._put(oh_addr, header)
		struct.pack_into("<QQ", self.buf, oh_addr + patch, data_addr, len(raw))
		return oh_addr
