"""CPU: the HDF5 + LZF writer of avvad/h5min.py against files the reference itself produced through
h5py (tests/golden/h5/*.h5, copied verbatim from data/subset) and against its own reader.

Reference: scripts/create_video_train_files_upsampled.py:244-310,373-385 and create_audio_train_files.py:182-193 write
X (67,67,T) / Y (y_dim,T) / statistics (513,1) and (1,1), float32, compression='lzf', chunks=None (auto)."""
import os
import struct

import numpy as np
import pytest

from avvad import h5min
from avvad.h5min import H5File, H5Writer, guess_chunk, write_h5
from util import GOLDEN


def _messages(path, name):
    h = H5File(path)
    return {t: bytes(d) for t, d in h._messages(h.datasets["/" + name])}


def test_auto_chunk_rule_matches_the_shipped_files():
    # shapes the reference CREATES the datasets with (then resizes): (67,67,0), (1,0), (513,0); statistics (513,1), (1,1)
    assert guess_chunk((67, 67, 0), (67, 67, None), 4) == (9, 9, 128)      # *_upsampled.h5 /X
    assert guess_chunk((1, 0), (1, None), 4) == (1, 1024)                  # *_vad_labels.h5 /Y
    assert guess_chunk((513, 0), (513, None), 4) == (33, 128)              # *_ibm_labels.h5 /Y
    assert guess_chunk((513, 1), (513, 1), 4) == (513, 1)                  # power_spec_statistics.h5
    assert guess_chunk((1, 1), (1, 1), 4) == (1, 1)                        # ntcd_timit_statistics.h5


@pytest.mark.parametrize("fname,key,kw", [
    ("sa1_vad_labels.h5", "Y", dict(maxshape=(1, None), creation_shape=(1, 0))),
    ("ntcd_timit_statistics.h5", "X_train_mean", dict(maxshape=(1, 1))),
])
def test_rewritten_file_carries_the_same_header_messages_and_data(tmp_path, fname, key, kw):
    src = os.path.join(GOLDEN, "h5", fname)
    ref = H5File(src)
    data = {k.lstrip("/"): ref[k] for k in ref.keys()}
    out = str(tmp_path / fname)
    write_h5(out, {k: (v, kw) for k, v in data.items()})
    got = H5File(out)
    assert got.keys() == ref.keys()
    for k in data:
        assert got[k].dtype == data[k].dtype and np.array_equal(got[k], data[k])
    a, b = _messages(src, key), _messages(out, key)
    for t in (0x01, 0x03, 0x05, 0x0B):            # dataspace (+max dims), datatype, fill value, LZF filter pipeline
        assert a[t] == b[t], hex(t)
    # layout: same version/class/rank and chunk dimensions (the B-tree address differs)
    assert a[0x08][:3] == b[0x08][:3] and a[0x08][11:] == b[0x08][11:]
    # superblock: same versions, offset/length sizes and group K values
    assert open(src, "rb").read(24) == open(out, "rb").read(24)


def test_lzf_encoder_reproduces_h5py_chunks_byte_for_byte():
    """The first chunks of the reference's test/34M/sa1_upsampled.h5 in write order: decode, re-encode with the hash
    table carried from chunk to chunk (as h5py's build of liblzf does), compare the stored bytes."""
    g = np.load(os.path.join(GOLDEN, "golden_lzf_chunks.npz"))
    cb = int(np.prod(g["chunk_dims"]))
    table = np.zeros(1 << 17, dtype=np.uint32)
    n = same = 0
    while f"chunk{n}" in g:
        stored = g[f"chunk{n}"].tobytes()
        raw = h5min.lzf_decompress(stored, cb)
        assert raw == h5min.lzf_decompress_py(stored, cb)       # native and pure-Python decoders agree
        again = h5min.lzf_compress(raw, table)
        assert h5min.lzf_decompress(again, cb) == raw
        same += again == stored
        n += 1
    assert n >= 8 and same == n, (same, n)


@pytest.mark.parametrize("shape,dtype,kw", [
    ((67, 67, 317), np.float32, dict(maxshape=(67, 67, None), creation_shape=(67, 67, 0))),   # 192 chunks: 2-level B-tree
    ((513, 300), np.float32, dict(maxshape=(513, None), creation_shape=(513, 0))),
    ((1, 1), np.float32, dict()),
    ((5, 7), np.float64, dict(chunks=(2, 3))),
    ((1000,), np.int32, dict(chunks=(64,))),
    ((3, 40), np.uint8, dict(compression=None, chunks=(3, 16))),
    ((67, 67, 0), np.float32, dict(maxshape=(67, 67, None))),                                  # created, never filled
])
def test_round_trip_through_the_reader(tmp_path, shape, dtype, kw):
    rng = np.random.default_rng(1)
    if np.issubdtype(dtype, np.floating):
        x = np.round(rng.standard_normal(shape) * 40).astype(dtype)     # compressible, like pixel / label data
    else:
        x = rng.integers(0, 100, shape).astype(dtype)
    out = str(tmp_path / "t.h5")
    with H5Writer(out) as f:
        f.create_dataset("X", x, **kw)
        f.create_dataset("Y", np.ones((1, max(1, shape[-1])), np.float32), maxshape=(1, None))
    h = H5File(out)
    assert h.keys() == ["/X", "/Y"]
    assert h["X"].shape == x.shape and h["X"].dtype == x.dtype and np.array_equal(h["X"], x)
    assert np.array_equal(h["Y"], np.ones((1, max(1, shape[-1])), np.float32))


def test_incompressible_chunks_are_stored_raw_with_the_filter_masked(tmp_path):
    x = np.random.default_rng(0).random((64, 64)).astype(np.float32)     # random mantissas: LZF cannot shrink them
    out = str(tmp_path / "r.h5")
    write_h5(out, {"X": (x, dict(chunks=(64, 64)))})
    h = H5File(out)
    assert np.array_equal(h["X"], x)
    layout = [d for t, d in h._messages(h.datasets["/X"]) if t == 0x08][0]
    btree = struct.unpack_from("<Q", layout, 3)[0]
    nbytes, fmask = struct.unpack_from("<II", h.buf, btree + 24)
    assert nbytes == x.nbytes and fmask == 1


def test_wav_writer_round_trip(tmp_path):
    x = (np.random.default_rng(2).standard_normal(4000) * 3000).astype(np.int16)
    p = str(tmp_path / "a.wav")
    h5min.write_wav_int16(p, x)
    y, fs = h5min.read_wav_int16(p)
    assert fs == 16000 and np.array_equal(x, y)
