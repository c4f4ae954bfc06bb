"""Minimal pure-Python HDF5 reader (test/tooling infrastructure, not product code).

h5py is not installed in the build image, but the reference's golden artefacts under
``data/subset`` are HDF5: processed ``*.h5`` (LZF-compressed chunks, h5py filter id 32000) and
MATLAB v7.3 ``*.mat`` (HDF5 behind a 512-byte user block, deflate chunks).  All of them are
"old style" files: superblock v0, v1 object headers, v1 B-trees, symbol-table groups, layout
message v3.  This reader supports exactly that subset and is used by
``tools/make_golden.py`` to turn the reference's own files into small fixtures under
``tests/golden/``.

Only little-endian IEEE floats and fixed-point integers are decoded.
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, List, Tuple

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"


def lzf_decompress(src: bytes, out_len: int) -> bytes:
    """liblzf decoder: ctrl<32 -> ctrl+1 literals; else back-reference (len=ctrl>>5 [+next], off)."""
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
