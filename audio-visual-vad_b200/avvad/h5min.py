"""Minimal HDF5 reader AND writer for the reference's on-disk formats -- h5py is not installed here.

h5py is not installed in the build image, but the reference's golden artefacts under
``data/subset`` are HDF5: processed ``*.h5`` (LZF-compressed chunks, h5py filter id 32000) and
MATLAB v7.3 ``*.mat`` (HDF5 behind a 512-byte user block, deflate chunks).  All of them are
"old style" files: superblock v0, v1 object headers, v1 B-trees, symbol-table groups, layout
message v3.  This reader supports exactly that subset and is used by
``tools/make_golden.py`` to turn the reference's own files into small fixtures under
``tests/golden/``.

Only little-endian IEEE floats and fixed-point integers are decoded.

The writer (:class:`H5Writer`) produces the same kind of file the reference's preparation scripts produce through h5py
(scripts/create_video_train_files_upsampled.py:244-310,373-385; create_audio_train_files.py:182-193): superblock v0, a
root symbol-table group, one v1 object header per dataset (dataspace with maximum dimensions, IEEE datatype, fill value,
LZF filter pipeline -- filter id 32000 --, chunked layout v3 indexed by a v1 B-tree, modification time), chunk shapes
from h5py's auto-chunking rule applied to the shape the dataset is CREATED with (the scripts create X as (67,67,0) and
Y as (y_dim,0), then resize).  The LZF codec is libavvad's host function (csrc/hostio.cu) when the library is built
(byte-identical with h5py's on > 99.8 % of the chunks of the shipped files, identical decoded data on all), with a
pure-Python decoder as the import-time-safe fallback for reading.
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, List, Tuple

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"


def _native():
    """libavvad's host codec (None when the shared library has not been built)."""
    try:
        from . import lib as L
        return L.lib()
    except Exception:
        return None


def lzf_decompress(src: bytes, out_len: int) -> bytes:
    """liblzf decoder: ctrl<32 -> ctrl+1 literals; else back-reference (len=ctrl>>5 [+next], off)."""
    l = _native()
    if l is not None:
        import ctypes
        out = (ctypes.c_uint8 * out_len)()
        n = l.avvad_lzf_decompress(bytes(src), len(src), out, out_len)
        if n != out_len:
            raise ValueError(f"lzf: produced {n} bytes, expected {out_len}")
        return bytes(out)
    return lzf_decompress_py(src, out_len)


def lzf_compress(raw: bytes, table=None, hlog=17):
    """LZF-compress one chunk with libavvad's host codec; None when the result would not be smaller (the chunk is then
    stored raw with the filter marked as skipped, as HDF5 does).  `table`: numpy uint32[2**hlog] carried across chunks."""
    import ctypes
    l = _native()
    if l is None:
        raise RuntimeError("writing LZF chunks needs libavvad (make -C audio-visual-vad_b200/csrc)")
    out = (ctypes.c_uint8 * len(raw))()
    n = l.avvad_lzf_compress(bytes(raw), len(raw), out, len(raw), hlog, None if table is None else table.ctypes.data)
    return bytes(out[:n]) if n > 0 else None


def lzf_decompress_py(src: bytes, out_len: int) -> bytes:
    out = bytearray(out_len)
    ip, op, n = 0, 0, len(src)
    while ip < n:
        ctrl = src[ip]
        ip += 1
        if ctrl < 32:
            run = ctrl + 1
            out[op:op + run] = src[ip:ip + run]
            ip += run
            op += run
        else:
            length = ctrl >> 5
            if length == 7:
                length += src[ip]
                ip += 1
            off = ((ctrl & 31) << 8) + src[ip] + 1
            ip += 1
            length += 2
            ref = op - off
            if ref < 0:
                raise ValueError("lzf: bad back-reference")
            if off >= length:
                out[op:op + length] = out[ref:ref + length]
            else:  # overlapping copy, byte by byte semantics
                for i in range(length):
                    out[op + i] = out[ref + i]
            op += length
    if op != out_len:
        raise ValueError(f"lzf: produced {op} bytes, expected {out_len}")
    return bytes(out)


class H5File:
    def __init__(self, path: str):
        with open(path, "rb") as f:
            self.buf = f.read()
        base = None
        for off in (0, 512, 1024, 2048):
            if self.buf[off:off + 8] == _SIG:
                base = off
                break
        if base is None:
            raise ValueError("not an HDF5 file: " + path)
        b = self.buf
        ver = b[base + 8]
        if ver != 0:
            raise NotImplementedError(f"superblock version {ver}")
        self.O = b[base + 13]
        self.L = b[base + 14]
        assert self.O == 8 and self.L == 8
        p = base + 24  # after group K values and consistency flags
        # base address, free-space, eof, driver info; every file address is relative to base_addr
        base_addr = struct.unpack_from("<Q", b, p)[0]
        p += 4 * self.O
        if base_addr:
            self.buf = b = self.buf[base_addr:]
            p -= base_addr
        # root symbol table entry
        _name_off, root_hdr, cache_type = struct.unpack_from("<QQI", b, p)
        self.root_hdr = root_hdr
        self.datasets: Dict[str, int] = {}
        self._walk_group(root_hdr, "")

    # ---- low-level structures -------------------------------------------------------------
    def _messages(self, addr: int) -> List[Tuple[int, bytes]]:
        b = self.buf
        ver, _, nmsg, _ref, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise NotImplementedError(f"object header version {ver}")
        msgs: List[Tuple[int, bytes]] = []
        blocks = [(addr + 16, hsize)]
        while blocks and len(msgs) < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(msgs) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, p)
                data = b[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x10:
                    coff, clen = struct.unpack_from("<QQ", data, 0)
                    blocks.append((coff, clen))
                msgs.append((mtype, data))
        return msgs

    def _heap_name(self, heap_addr: int, off: int) -> str:
        b = self.buf
        assert b[heap_addr:heap_addr + 4] == b"HEAP"
        data_addr = struct.unpack_from("<Q", b, heap_addr + 8 + 2 * self.L)[0]
        s = data_addr + off
        e = b.index(b"\x00", s)
        return b[s:e].decode()

    def _walk_btree_group(self, addr: int, heap: int, prefix: str):
        b = self.buf
        assert b[addr:addr + 4] == b"TREE", "bad group btree"
        ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
        assert ntype == 0
        p = addr + 8 + 2 * self.O
        p += self.L  # key 0
        for _ in range(used):
            child = struct.unpack_from("<Q", b, p)[0]
            p += self.O + self.L
            if level > 0:
                self._walk_btree_group(child, heap, prefix)
            else:
                self._walk_snod(child, heap, prefix)

    def _walk_snod(self, addr: int, heap: int, prefix: str):
        b = self.buf
        assert b[addr:addr + 4] == b"SNOD"
        nsym = struct.unpack_from("<H", b, addr + 6)[0]
        p = addr + 8
        for _ in range(nsym):
            name_off, hdr, cache = struct.unpack_from("<QQI", b, p)
            p += 40
            name = self._heap_name(heap, name_off)
            self._visit(hdr, prefix + "/" + name)

    def _visit(self, hdr: int, path: str):
        msgs = self._messages(hdr)
        types = {t for t, _ in msgs}
        if 0x11 in types:
            self._walk_group(hdr, path)
        elif 0x08 in types:
            self.datasets[path] = hdr

    def _walk_group(self, hdr: int, prefix: str):
        for t, d in self._messages(hdr):
            if t == 0x11:
                btree, heap = struct.unpack_from("<QQ", d, 0)
                self._walk_btree_group(btree, heap, prefix)

    # ---- dataset decoding --------------------------------------------------------------------
    def keys(self):
        return sorted(self.datasets)

    def __getitem__(self, name: str) -> np.ndarray:
        if not name.startswith("/"):
            name = "/" + name
        hdr = self.datasets[name]
        shape = dtype = layout = None
        filters: List[Tuple[int, List[int]]] = []
        for t, d in self._messages(hdr):
            if t == 0x01:
                ver, rank, flags = struct.unpack_from("<BBB", d, 0)
                assert ver == 1
                shape = struct.unpack_from("<" + "Q" * rank, d, 8)
            elif t == 0x03:
                cls = d[0] & 0x0F
                size = struct.unpack_from("<I", d, 4)[0]
                bits0 = d[1]
                assert (bits0 & 1) == 0, "big-endian not supported"
                if cls == 1:
                    dtype = np.dtype("<f%d" % size)
                elif cls == 0:
                    signed = (bits0 >> 3) & 1
                    dtype = np.dtype("<%s%d" % ("i" if signed else "u", size))
                else:
                    raise NotImplementedError(f"datatype class {cls}")
            elif t == 0x08:
                layout = d
            elif t == 0x0B:
                ver, nf = struct.unpack_from("<BB", d, 0)
                assert ver == 1
                p = 8
                for _ in range(nf):
                    fid, nlen, _fl, ncd = struct.unpack_from("<HHHH", d, p)
                    p += 8
                    p += (nlen + 7) // 8 * 8
                    cd = list(struct.unpack_from("<" + "I" * ncd, d, p))
                    p += 4 * ncd
                    if ncd % 2:
                        p += 4
                    filters.append((fid, cd))
        assert shape is not None and dtype is not None and layout is not None
        lver, lclass = layout[0], layout[1]
        assert lver == 3, f"layout version {lver}"
        if lclass == 1:  # contiguous
            addr, size = struct.unpack_from("<QQ", layout, 2)
            n = int(np.prod(shape)) if len(shape) else 1
            return np.frombuffer(self.buf, dtype=dtype, count=n, offset=addr).reshape(shape).copy()
        if lclass == 0:  # compact
            size = struct.unpack_from("<H", layout, 2)[0]
            return np.frombuffer(layout[4:4 + size], dtype=dtype).reshape(shape).copy()
        assert lclass == 2
        ndim = layout[2]
        btree = struct.unpack_from("<Q", layout, 3)[0]
        cdims = struct.unpack_from("<" + "I" * ndim, layout, 3 + self.O)
        chunk_shape = cdims[:-1]
        rank = ndim - 1
        assert rank == len(shape)
        out = np.zeros(shape, dtype=dtype)
        chunk_bytes = int(np.prod(chunk_shape)) * dtype.itemsize
        self._read_chunks(btree, rank, chunk_shape, chunk_bytes, filters, dtype, out)
        return out

    def _read_chunks(self, addr, rank, chunk_shape, chunk_bytes, filters, dtype, out):
        b = self.buf
        assert b[addr:addr + 4] == b"TREE", "bad chunk btree"
        ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
        assert ntype == 1
        p = addr + 8 + 2 * self.O
        keysize = 8 + 8 * (rank + 1)
        for _ in range(used):
            nbytes, fmask = struct.unpack_from("<II", b, p)
            offs = struct.unpack_from("<" + "Q" * (rank + 1), b, p + 8)
            child = struct.unpack_from("<Q", b, p + keysize)[0]
            p += keysize + self.O
            if level > 0:
                self._read_chunks(child, rank, chunk_shape, chunk_bytes, filters, dtype, out)
                continue
            raw = b[child:child + nbytes]
            for i in reversed(range(len(filters))):
                if fmask & (1 << i):
                    continue
                fid, cd = filters[i]
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 32000:
                    raw = lzf_decompress(raw, chunk_bytes)
                elif fid == 2:  # shuffle
                    a = np.frombuffer(raw, dtype=np.uint8).reshape(dtype.itemsize, -1)
                    raw = a.T.tobytes()
                else:
                    raise NotImplementedError(f"filter {fid}")
            chunk = np.frombuffer(raw, dtype=dtype, count=int(np.prod(chunk_shape))).reshape(chunk_shape)
            sl_out, sl_in = [], []
            for d in range(rank):
                s = offs[d]
                e = min(s + chunk_shape[d], out.shape[d])
                sl_out.append(slice(s, e))
                sl_in.append(slice(0, e - s))
            out[tuple(sl_out)] = chunk[tuple(sl_in)]


def read_wav_int16(path: str):
    """RIFF/PCM 16-bit reader returning (int16 array (n,), sample_rate). Mono files only."""
    import wave

    with wave.open(path, "rb") as w:
        assert w.getsampwidth() == 2 and w.getnchannels() == 1
        fs = w.getframerate()
        data = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").copy()
    return data, fs


# ---------------------------------------------------------------------------------------------
# writer
# ---------------------------------------------------------------------------------------------
_UNDEF = 0xFFFFFFFFFFFFFFFF
_CHUNK_BASE, _CHUNK_MIN, _CHUNK_MAX = 16 * 1024, 8 * 1024, 1024 * 1024
_GROUP_LEAF_K, _GROUP_INTERNAL_K, _CHUNK_K = 4, 16, 32


def guess_chunk(shape, maxshape, typesize):
    """h5py's auto-chunking rule (dimensions of size 0 count as 1024; halve the axes round-robin until the chunk is
    within 50 % of a target between 8 KiB and 1 MiB that grows with log10 of the dataset size)."""
    chunks = np.array([float(s) if s != 0 else 1024.0 for s in shape], dtype="=f8")
    ndims = len(shape)
    if ndims == 0:
        raise ValueError("chunks not allowed for scalar datasets")
    dset_size = np.prod(chunks) * typesize
    target = _CHUNK_BASE * (2 ** np.log10(dset_size / (1024.0 * 1024)))
    target = min(max(target, _CHUNK_MIN), _CHUNK_MAX)
    idx = 0
    while True:
        cb = np.prod(chunks) * typesize
        if (cb < target or abs(cb - target) / target < 0.5) and cb < _CHUNK_MAX:
            break
        if np.prod(chunks) == 1:
            break
        chunks[idx % ndims] = np.ceil(chunks[idx % ndims] / 2.0)
        idx += 1
    return tuple(int(x) for x in chunks)


def _datatype_message(dt: np.dtype) -> bytes:
    dt = np.dtype(dt).newbyteorder("<")
    if dt.kind == "f" and dt.itemsize == 4:
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 0x1F, 0x00, 4, 0, 32, 23, 8, 0, 23, 127)
    if dt.kind == "f" and dt.itemsize == 8:
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 0x3F, 0x00, 8, 0, 64, 52, 11, 0, 52, 1023)
    if dt.kind in "iu":
        return struct.pack("<BBBBIHH", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize, 0, 8 * dt.itemsize) + b"\0" * 4
    raise NotImplementedError(f"datatype {dt}")


def _msg(mtype: int, data: bytes, flags=0) -> bytes:
    data = data + b"\0" * (-len(data) % 8)
    return struct.pack("<HHBBBB", mtype, len(data), flags, 0, 0, 0) + data


class H5Writer:
    """with H5Writer(path) as f: f.create_dataset('X', array, maxshape=(67, 67, None), creation_shape=(67, 67, 0))"""

    def __init__(self, path: str, mtime: int = 0):
        self.path, self.mtime = path, int(mtime)
        self.items = []

    def __enter__(self):
        return self

    def __exit__(self, et, ev, tb):
        if et is None:
            self.close()
        return False

    def create_dataset(self, name, data, maxshape=None, chunks=None, compression="lzf", creation_shape=None):
        """`chunks=None` + compression -> h5py's auto-chunking of `creation_shape` (default: data.shape), i.e. what
        `create_dataset(shape=creation_shape, maxshape=..., chunks=None, compression='lzf')` followed by resize + write
        leaves on disk."""
        data = np.ascontiguousarray(data)
        if data.dtype.byteorder == ">":
            data = data.astype(data.dtype.newbyteorder("<"))
        if data.ndim == 0:
            raise NotImplementedError("scalar datasets")
        if compression not in (None, "lzf"):
            raise NotImplementedError(f"compression {compression!r}")
        maxshape = tuple(data.shape) if maxshape is None else tuple(maxshape)
        if chunks is None:
            chunks = guess_chunk(tuple(creation_shape) if creation_shape is not None else data.shape,
                                 maxshape, data.dtype.itemsize)
        assert len(chunks) == data.ndim and len(maxshape) == data.ndim
        if any(n == name for n, *_ in self.items):
            raise ValueError(f"dataset {name!r} exists")
        self.items.append((name, data, maxshape, tuple(int(c) for c in chunks), compression))

    # ---- file assembly ---------------------------------------------------------------------
    def close(self):
        if len(self.items) > 2 * _GROUP_LEAF_K:
            raise NotImplementedError("more datasets than one symbol-table node holds")
        items = sorted(self.items, key=lambda it: it[0])   # symbol-table nodes are sorted by name
        out = bytearray()

        def alloc(n, align=8):
            pad = -len(out) % align
            out.extend(b"\0" * pad)
            addr = len(out)
            out.extend(b"\0" * n)
            return addr

        sb = alloc(96)
        root_hdr = alloc(16 + 8 + 16)
        btree = alloc(24 + (2 * _GROUP_INTERNAL_K + 1) * 8 + 2 * _GROUP_INTERNAL_K * 8)
        # local heap: offset 0 holds the empty string; names are NUL-terminated and 8-byte aligned
        heap_data = bytearray(b"\0" * 8)
        name_off = {}
        for name, *_ in items:
            name_off[name] = len(heap_data)
            nb = name.encode() + b"\0"
            heap_data.extend(nb + b"\0" * (-len(nb) % 8))
        heap_hdr = alloc(32)
        heap_seg = alloc(len(heap_data))
        out[heap_seg:heap_seg + len(heap_data)] = heap_data
        snod = alloc(8 + 2 * _GROUP_LEAF_K * 40)

        table = np.zeros(1 << 17, dtype=np.uint32)   # LZF hash table carried across the chunks of the file
        hdr_addr = {}
        for name, data, maxshape, chunks, compression in items:
            rank, esz = data.ndim, data.dtype.itemsize
            # ---- chunks, in row-major chunk order ----
            grid = [(-(-s // c)) for s, c in zip(data.shape, chunks)]
            entries = []
            cb = int(np.prod(chunks)) * esz
            for idx in np.ndindex(*grid) if all(g > 0 for g in grid) else []:
                offs = tuple(i * c for i, c in zip(idx, chunks))
                block = np.zeros(chunks, dtype=data.dtype)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunks, data.shape))
                block[tuple(slice(0, x.stop - x.start) for x in sl)] = data[sl]
                raw = block.tobytes()
                comp = lzf_compress(raw, table) if compression == "lzf" else None
                payload, mask = (comp, 0) if comp is not None else (raw, 1 if compression == "lzf" else 0)
                a = alloc(len(payload), 1)
                out[a:a + len(payload)] = payload
                entries.append((len(payload), mask, offs, a))
            # ---- chunk B-tree (v1, node type 1), as many levels as needed ----
            keysize = 8 + 8 * (rank + 1)
            node_bytes = 24 + (2 * _CHUNK_K + 1) * keysize + 2 * _CHUNK_K * 8
            end_key = (0, 0, tuple(g * c for g, c in zip(grid, chunks)))   # one past the last chunk

            def key_bytes(size, mask, offs):
                return struct.pack("<II", size, mask) + struct.pack("<" + "Q" * (rank + 1), *offs, 0)

            def write_level(children, level):
                """children: list of (first key fields, address).  Returns the list of nodes of this level."""
                nodes = []
                groups = [children[i:i + 2 * _CHUNK_K] for i in range(0, len(children), 2 * _CHUNK_K)] or [[]]
                addrs = [alloc(node_bytes) for _ in groups]
                for gi, grp in enumerate(groups):
                    left = addrs[gi - 1] if gi > 0 else _UNDEF
                    right = addrs[gi + 1] if gi + 1 < len(groups) else _UNDEF
                    b = bytearray(b"TREE" + struct.pack("<BBHQQ", 1, level, len(grp), left, right))
                    for key, child in grp:
                        b += key_bytes(*key) + struct.pack("<Q", child)
                    nxt = groups[gi + 1][0][0] if gi + 1 < len(groups) else end_key
                    b += key_bytes(*nxt)
                    out[addrs[gi]:addrs[gi] + len(b)] = b
                    nodes.append((grp[0][0] if grp else end_key, addrs[gi]))
                return nodes

            level = 0
            nodes = write_level([((sz, mk, offs), a) for sz, mk, offs, a in entries], 0)
            while len(nodes) > 1:
                level += 1
                nodes = write_level(nodes, level)
            chunk_btree = nodes[0][1]
            # ---- object header ----
            dims = struct.pack("<" + "Q" * rank, *data.shape)
            maxd = struct.pack("<" + "Q" * rank, *[_UNDEF if m is None else int(m) for m in maxshape])
            msgs = _msg(0x01, struct.pack("<BBBBI", 1, rank, 1, 0, 0) + dims + maxd)
            msgs += _msg(0x03, _datatype_message(data.dtype), flags=1)
            msgs += _msg(0x05, struct.pack("<BBBBI", 2, 3, 0, 1, 0))
            if compression == "lzf":
                msgs += _msg(0x0B, struct.pack("<BB6x", 1, 1) + struct.pack("<HHHH", 32000, 8, 1, 3) + b"lzf\0\0\0\0\0" +
                             struct.pack("<III", 4, 0x0105, cb) + b"\0" * 4, flags=1)
            msgs += _msg(0x08, struct.pack("<BBB", 3, 2, rank + 1) + struct.pack("<Q", chunk_btree) +
                         struct.pack("<" + "I" * (rank + 1), *chunks, esz))
            msgs += _msg(0x12, struct.pack("<B3xI", 1, self.mtime))
            nmsg = 6 if compression == "lzf" else 5
            h = alloc(16 + len(msgs))
            out[h:h + 16] = struct.pack("<BBHII4x", 1, 0, nmsg, 1, len(msgs))
            out[h + 16:h + 16 + len(msgs)] = msgs
            hdr_addr[name] = h

        # ---- group structures ----
        b = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(items)))
        for name, *_ in items:
            b += struct.pack("<QQII16x", name_off[name], hdr_addr[name], 0, 0)
        out[snod:snod + len(b)] = b
        out[heap_hdr:heap_hdr + 32] = b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), 1, heap_seg)
        last = name_off[items[-1][0]] if items else 0
        nb = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if items else 0, _UNDEF, _UNDEF))
        nb += struct.pack("<Q", 0) + struct.pack("<Q", snod) + struct.pack("<Q", last)
        out[btree:btree + len(nb)] = nb
        out[root_hdr:root_hdr + 40] = struct.pack("<BBHII4x", 1, 0, 1, 1, 24) + _msg(0x11, struct.pack("<QQ", btree, heap_hdr))
        eof = len(out)
        out[sb:sb + 96] = (_SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, _GROUP_LEAF_K, _GROUP_INTERNAL_K, 0) +
                           struct.pack("<QQQQ", 0, _UNDEF, eof, _UNDEF) +
                           struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", btree, heap_hdr))
        with open(self.path, "wb") as fh:
            fh.write(bytes(out))


def write_h5(path, datasets, mtime=0):
    """datasets: {name: array} or {name: (array, dict(maxshape=..., creation_shape=..., chunks=..., compression=...))}."""
    with H5Writer(path, mtime) as f:
        for name, v in datasets.items():
            if isinstance(v, tuple):
                f.create_dataset(name, v[0], **v[1])
            else:
                f.create_dataset(name, v)


def write_wav_int16(path: str, data, fs=16000):
    """Mono 16-bit RIFF/PCM writer (the corpus format)."""
    import wave

    data = np.asarray(data)
    if data.dtype != np.int16:
        data = np.clip(np.rint(data * 32768.0), -32768, 32767).astype(np.int16)
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(int(fs))
        w.writeframes(data.astype("<i2").tobytes())
