"""LightSparseDataFrame / read_lsdf behave like the reference's (sparse_utils.py) --
checked against a file written by the reference's own writer (tests/golden/lsdf_small.npz)."""
import json
import os

import numpy as np
import scipy.sparse

from conftest import GOLDEN
from pangenomix_b200 import sparse_utils as su
from pangenomix_b200 import synth


def _small():
    return su.read_lsdf(os.path.join(GOLDEN, "lsdf_small.npz"))


def test_read_reference_written_file():
    lsdf = _small()
    coo = synth.bernoulli_matrix(800, 50, 450, seed=7)
    assert lsdf.shape == (800, 50)
    assert (lsdf.data.tocsr() != coo.tocsr()).nnz == 0
    assert lsdf.data.format == "coo" and lsdf.data.data.dtype == np.int64
    assert lsdf.index[0] == "T_C0" and lsdf.columns[-1] == "genome49"
    assert lsdf.index_map["T_C17"] == 17 and lsdf.column_map["genome3"] == 3
    manifest = json.load(open(os.path.join(GOLDEN, "MANIFEST.json")))
    assert [int(v) for v in lsdf.sum(axis="index")[:8]] == manifest["lsdf_small_row_sums_head"]
    assert [int(v) for v in lsdf.sum(axis=1)[:8]] == manifest["lsdf_small_col_sums_head"]


def test_round_trip_and_file_format(tmp_path):
    lsdf = _small()
    path = str(tmp_path / "t.npz")
    lsdf.to_npz(path)
    raw = np.load(path)
    assert set(raw.files) == {"row", "col", "data", "shape", "format"}
    assert raw["format"].item() in (b"coo", "coo")
    lines = open(path + ".labels.txt").read().split("\n")
    assert lines[0] == "T_C0" and lines[800] == "genome0" and len(lines) == 851
    back = su.read_lsdf(path)
    assert (back.data.tocsr() != lsdf.data.tocsr()).nnz == 0
    assert list(back.index) == list(lsdf.index) and list(back.columns) == list(lsdf.columns)


def test_slicing_transpose_drop_empty(capsys):
    lsdf = _small()
    dense = lsdf.values
    sub = lsdf.islice([5, 1, 7], [3, 2])
    assert np.array_equal(sub.values, dense[[5, 1, 7]][:, [3, 2]])
    assert list(sub.index) == ["T_C5", "T_C1", "T_C7"] and list(sub.columns) == ["genome3", "genome2"]
    assert np.array_equal(lsdf.labelslice(columns=["genome9"]).values, dense[:, [9]])
    assert np.array_equal(lsdf.iloc[[2, 3]].values, dense[[2, 3]])
    assert np.array_equal(lsdf.transpose().values, dense.T)
    assert lsdf.islice() is None and "No indices or columns selected" in capsys.readouterr().out
    x = scipy.sparse.coo_matrix(np.array([[0, 1, 0], [0, 0, 0], [1, 1, 0]]))
    t = su.LightSparseDataFrame(["a", "b", "c"], ["x", "y", "z"], x)
    assert list(t.drop_empty().index) == ["a", "c"]
    assert list(t.drop_empty(axis="columns").columns) == ["x", "y"]
    assert t.npoints == 3 and list(t.sp_index) == ["a", "b", "c"]


def test_constructor_diagnostics(capsys):
    x = scipy.sparse.coo_matrix(np.eye(2, dtype=np.int64))
    su.LightSparseDataFrame(["a"], ["x", "y", "z"], x)
    out = capsys.readouterr().out
    assert "ERROR: Index length does not match data" in out
    assert "ERROR: Column length does no match data" in out


def test_compress_rows_and_sparse_array_round_trip():
    x = np.array([[1, 0, 1], [0, 1, 0], [1, 0, 1], [0, 0, 0], [0, 1, 0]], dtype=np.int64)
    t = su.LightSparseDataFrame(list("abcde"), list("xyz"), scipy.sparse.coo_matrix(x))
    blocks, members = su.compress_rows(t)
    assert list(blocks.index) == ["B0", "B1", "B2"]
    assert np.array_equal(blocks.values, x[[0, 1, 3]])
    assert [list(m) for m in members] == [["a", "c"], ["b", "e"], ["d"]]
    frame = t.to_sparse_arrays()
    assert frame.shape == (5, 3) and list(frame.index) == list("abcde")
    back = su.sparse_arrays_to_lsdf(frame)
    assert np.array_equal(np.nan_to_num(back.values), x.astype(float))


def test_threaded_npz_loader_equals_scipy(tmp_path):
    """read_lsdf inflates row / col / data of a COO archive on three threads: same matrix, same dtypes,
    writable arrays as scipy.sparse.load_npz gives; other archives are left to scipy."""
    ref_file = os.path.join(GOLDEN, "lsdf_small.npz")              # written by the reference's to_npz
    for path in (ref_file,):
        fast, slow = su._load_npz_coo(path), scipy.sparse.load_npz(path)
        assert fast is not None and fast.format == "coo" and fast.shape == slow.shape
        for name in ("row", "col", "data"):
            a, b = getattr(fast, name), getattr(slow, name)
            assert a.dtype == b.dtype and np.array_equal(a, b) and a.flags.writeable
    coo = synth.bernoulli_matrix(300, 20, 100, seed=2)
    stored = str(tmp_path / "stored.npz")
    scipy.sparse.save_npz(stored, coo, compressed=False)
    assert (su._load_npz_coo(stored).tocsr() != coo.tocsr()).nnz == 0
    as_csr = str(tmp_path / "csr.npz")
    scipy.sparse.save_npz(as_csr, coo.tocsr())
    assert su._load_npz_coo(as_csr) is None                         # not a COO archive: scipy's job
    labels = "\n".join(["r%d" % i for i in range(300)] + ["c%d" % i for i in range(20)])
    open(as_csr + ".labels.txt", "w").write(labels)
    assert su.read_lsdf(as_csr).shape == (300, 20)
    # a damaged member is not silently accepted
    blob = bytearray(open(ref_file, "rb").read())
    import zipfile
    info = {i.filename: i for i in zipfile.ZipFile(ref_file).infolist()}["row.npy"]
    blob[info.header_offset + 30 + len("row.npy") + info.compress_size // 2] ^= 0xFF
    broken = str(tmp_path / "broken.npz")
    open(broken, "wb").write(bytes(blob))
    assert su._load_npz_coo(broken) is None
    import pytest
    with pytest.raises(Exception):
        scipy.sparse.load_npz(broken)
