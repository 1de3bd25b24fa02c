"""CPU: the built-in HDF5 reader / writer (summarizer_b200/utils/hdf5.py, SURVEY.md §8f NEXT-3) — round trips of the
dataset schema (datasets/README.md:5-42) and of the ``<split>_preds.h5`` layout (models/__init__.py:149-177), a
specification-level walk of the bytes the writer emits (what libhdf5's loaders check), and a hand-assembled file with
the reader-only features (continuation block, compact and chunked + gzip + shuffle layouts, big-endian data,
variable-length strings)."""
import struct
import zlib

import numpy as np
import pytest

from summarizer_b200 import synthetic
from summarizer_b200.utils import hdf5


def test_dataset_schema_round_trip(tmp_path):
    ds = synthetic.make_dataset("summe", 3)
    path = synthetic.write_dataset_h5(ds, str(tmp_path / "summarizer_dataset_summe_google_pool5.h5"))
    assert hdf5.check_file(path) == 1 + 3 * (1 + len(ds["video_1"]))
    with hdf5.File(path, "r") as f:
        assert sorted(f.keys()) == sorted(ds.keys()) and "video_2" in f and "video_9" not in f
        for key in ds.keys():
            g = f[key]
            for name in dict.keys(ds[key]):
                want = ds[key].raw(name)
                got = g[name][...]
                if isinstance(want, str):
                    assert bytes(got).rstrip(b"\0").decode() == want, (key, name)
                    continue
                assert np.asarray(got).dtype == np.asarray(want).dtype, (key, name)
                assert np.array_equal(got, want), (key, name)
            assert int(g["n_frames"][()]) == int(ds[key].raw("n_frames"))          # the Trainer's access patterns
            assert g["n_frame_per_seg"][...].tolist() == ds[key].raw("n_frame_per_seg").tolist()
            assert g["features"].shape == ds[key].raw("features").shape and g["features"].dtype == np.float32
            assert np.array_equal(f[f"{key}/picks"][:5], ds[key].raw("picks")[:5])
        with pytest.raises(KeyError):
            f["video_1/nope"]


def test_predictions_file_layout_and_types(tmp_path):
    path = str(tmp_path / "tvsum_splits.json_preds.h5")
    rng = np.random.default_rng(0)
    want = {}
    with hdf5.File(path, "w") as f:
        d = f.create_group("summarizer_dataset_tvsum_google_pool5.h5")
        for i in range(70):                                           # more links than one symbol-table node holds
            k = d.create_group(f"video_{i}")
            want[i] = dict(scores=rng.random(20).astype(np.float32), user_summary=(rng.random((3, 50)) > 0.5).astype(np.float32),
                           machine_summary=(rng.random(50) > 0.8).astype(np.float32), machine_scores=rng.random(50).astype(np.float32))
            for name, arr in want[i].items():
                k.create_dataset(name, data=arr)
        extra = f.create_group("types")
        for dt in (np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint32, np.float64):
            extra.create_dataset(np.dtype(dt).name, data=np.arange(-3, 4).astype(dt))
        extra.create_dataset("scalar", data=np.int32(4494))
        extra.create_dataset("empty", data=np.zeros((0, 4), np.float32))
        extra.create_dataset("flags", data=np.array([True, False, True]))
        with pytest.raises(ValueError):
            extra.create_dataset("scalar", data=1)
        with pytest.raises(TypeError):
            extra.create_dataset("cplx", data=np.zeros(2, np.complex64))
    n = hdf5.check_file(path)
    assert n == 1 + 1 + 70 * 5 + 1 + 10
    with hdf5.File(path, "r") as f:
        g = f["summarizer_dataset_tvsum_google_pool5.h5"]
        assert sorted(g.keys()) == sorted(f"video_{i}" for i in range(70)) and len(g) == 70
        for i in (0, 9, 10, 69):
            for name, arr in want[i].items():
                got = g[f"video_{i}"][name][...]
                assert got.dtype == arr.dtype and np.array_equal(got, arr)
        # summary.py:40-43
        assert np.array_equal(f["summarizer_dataset_tvsum_google_pool5.h5"]["video_7"]["machine_summary"][...], want[7]["machine_summary"])
        t = f["types"]
        for dt in (np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint32, np.float64):
            got = t[np.dtype(dt).name][...]
            assert got.dtype == np.dtype(dt) and np.array_equal(got, np.arange(-3, 4).astype(dt))
        assert t["scalar"][()] == 4494 and t["scalar"].shape == ()
        assert t["empty"][...].shape == (0, 4)
        assert t["flags"][...].tolist() == [1, 0, 1]


def test_not_hdf5_and_unsupported_files_raise(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file" * 8)
    with pytest.raises(OSError):
        hdf5.File(str(p), "r")
    sb = bytearray(hdf5.SIGNATURE + bytes(120))
    sb[8] = 2                                                         # superblock version 2 (libver='latest')
    p.write_bytes(bytes(sb))
    with pytest.raises(NotImplementedError):
        hdf5.File(str(p), "r")


# ---- a hand-assembled file with the structures only the READER has to understand --------------------------------
def _msg(t, body, flags=0):
    body += b"\0" * (-len(body) % 8)
    return struct.pack("<HHB3x", t, len(body), flags) + body


def _ohdr(msgs, n_total=None, chunk=None):
    chunk = b"".join(msgs) if chunk is None else chunk
    return struct.pack("<BBHII4x", 1, 0, n_total or len(msgs), 1, len(chunk)) + chunk


def test_reader_only_features(tmp_path):
    """compact layout, header continuation, chunked + shuffle + gzip, big-endian ints, variable-length string."""
    blob = bytearray(96)
    put = lambda b: (blob.extend(b"\0" * (-len(blob) % 8)), len(blob), blob.extend(b))[1]
    i32be = struct.pack("<BBBBI", 0x10, 0x09, 0, 0, 4) + struct.pack("<HH", 0, 32)          # big endian, signed
    f32 = hdf5.encode_datatype(np.float32)
    space1 = lambda n: struct.pack("<BBBB4x", 1, 1, 0, 0) + struct.pack("<Q", n)
    # (a) compact, big-endian
    a_data = np.arange(5, dtype=">i4").tobytes()
    a_hdr = put(_ohdr([_msg(1, space1(5)), _msg(3, i32be), _msg(8, struct.pack("<BBH", 3, 0, len(a_data)) + a_data)]))
    # (b) contiguous float data, layout message in a continuation block
    b_vals = np.linspace(0, 1, 7, dtype=np.float32)
    b_data = put(b_vals.tobytes())
    cont = put(_msg(8, struct.pack("<BBQQ", 3, 1, b_data, 28)) + _msg(0, b"\0" * 8))
    b_hdr = put(_ohdr([_msg(1, space1(7)), _msg(3, f32), _msg(0x10, struct.pack("<QQ", cont, 48))], n_total=5))
    # (c) chunked (4 per chunk, 10 elements), shuffle + deflate
    c_vals = (np.arange(10, dtype=np.float32) * 1.5)
    chunks = []
    for o in range(0, 10, 4):
        part = np.zeros(4, np.float32); part[: min(4, 10 - o)] = c_vals[o:o + 4]
        raw = np.frombuffer(part.tobytes(), np.uint8).reshape(4, 4).T.tobytes()            # shuffle
        z = zlib.compress(raw)
        chunks.append((o, put(z), len(z)))
    node = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(chunks), hdf5.UNDEF, hdf5.UNDEF)
    for o, addr, n in chunks:
        node += struct.pack("<IIQQ", n, 0, o, 0) + struct.pack("<Q", addr)
    node += struct.pack("<IIQQ", 0, 0, 12, 0)
    c_tree = put(node)
    filt = struct.pack("<BB6x", 1, 2) + struct.pack("<HHHH", 2, 0, 0, 1) + struct.pack("<II", 4, 0) \
        + struct.pack("<HHHH", 1, 0, 0, 1) + struct.pack("<II", 6, 0)
    c_hdr = put(_ohdr([_msg(1, space1(10)), _msg(3, f32), _msg(0xB, filt),
                       _msg(8, struct.pack("<BBBQII", 3, 2, 2, c_tree, 4, 4))]))
    # (d) scalar variable-length UTF-8 string through a global heap collection
    text = "Vidéo_42".encode("utf-8")
    obj = struct.pack("<HHIQ", 1, 1, 0, len(text)) + text + b"\0" * (-len(text) % 8)
    gcol_size = 16 + len(obj) + 16
    gcol = put(b"GCOL" + struct.pack("<B3xQ", 1, gcol_size) + obj + struct.pack("<HHIQ", 0, 0, 0, 16))
    vlen_t = struct.pack("<BBBBI", 0x19, 0x01, 0x01, 0, 16) + struct.pack("<BBBBI", 0x13, 0x10, 0, 0, 1)
    d_data = put(struct.pack("<IQI", len(text), gcol, 1))
    d_hdr = put(_ohdr([_msg(1, struct.pack("<BBBB4x", 1, 0, 0, 0)), _msg(3, vlen_t), _msg(8, struct.pack("<BBQQ", 3, 1, d_data, 16))]))
    # root group with the four links
    names = [b"a_compact", b"b_cont", b"c_chunked", b"d_name"]
    seg, offs = bytearray(8), []
    for n in names:
        offs.append(len(seg)); seg += n + b"\0" * (8 - len(n) % 8)
    heap = put(b"HEAP" + struct.pack("<B3xQQQ", 0, len(seg), 1, 0) + bytes(seg))
    blob[heap + 24:heap + 32] = struct.pack("<Q", heap + 32)
    snod = b"SNOD" + struct.pack("<BBH", 1, 0, 4)
    for o, h in zip(offs, (a_hdr, b_hdr, c_hdr, d_hdr)):
        snod += struct.pack("<QQII16x", o, h, 0, 0)
    snod_at = put(snod + b"\0" * (8 + 8 * 40 - len(snod)))
    tree = put(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, hdf5.UNDEF, hdf5.UNDEF) + struct.pack("<QQQ", 0, snod_at, offs[-1]) + b"\0" * 512)
    root = put(_ohdr([_msg(0x11, struct.pack("<QQ", tree, heap))]))
    sb = hdf5.SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0) + struct.pack("<QQQQ", 0, hdf5.UNDEF, len(blob), hdf5.UNDEF)
    sb += struct.pack("<QQII", 0, root, 0, 0) + bytes(16)             # root entry without cached symbol-table info
    blob[:96] = sb
    p = tmp_path / "hand.h5"
    p.write_bytes(bytes(blob))
    with hdf5.File(str(p), "r") as f:
        assert f.keys() == ["a_compact", "b_cont", "c_chunked", "d_name"]
        a = f["a_compact"][...]
        assert a.tolist() == [0, 1, 2, 3, 4] and a.dtype.byteorder in "=<|"
        assert np.array_equal(f["b_cont"][...], b_vals)
        assert np.array_equal(f["c_chunked"][...], c_vals)
        assert f["d_name"][()] == "Vidéo_42"


def test_summary_tool_reads_predictions(tmp_path):
    """summary.py:40-43: the mp4 tool picks ``machine_summary`` of one video out of the predictions file and maps kept
    frame i to ``%06d.jpg % (i + 1)`` (summary.py:14-16)."""
    from summarizer_b200 import summary as tool
    path = str(tmp_path / "summe_splits.json_preds.h5")
    ms = np.array([0, 1, 1, 0, 0, 1], np.float32)
    with hdf5.File(path, "w") as f:
        g = f.create_group("summe.h5").create_group("video_3")
        g.create_dataset("machine_summary", data=ms)
    got = tool.read_machine_summary(path, "summe.h5", "video_3")
    assert got.dtype == np.float32 and np.array_equal(got, ms)
    assert tool.kept_frame_names(got) == ["000002.jpg", "000003.jpg", "000006.jpg"]
