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


def _deflate(data, level=6, strategy=0, wbits=-15):
    import zlib
    c = zlib.compressobj(level, zlib.DEFLATED, wbits, 9, strategy)
    return c.compress(data) + c.flush()


def test_native_inflate_equals_zlib():
    """pgx_inflate_raw (csrc/pgx_inflate.cpp), the decoder behind read_lsdf: byte-identical to zlib on stored,
    fixed and dynamic blocks, short distances, long runs, multi-block streams; wrong sizes and damaged streams
    are refused or differ (the loader then checks the zip CRC and falls back to zlib)."""
    import zlib
    rng = np.random.RandomState(0)
    cases = [b"", b"a", b"abc" * 5, bytes(1000),
             bytes(rng.randint(0, 256, 70000, dtype=np.uint8)),                                   # stored blocks
             bytes(rng.randint(0, 4, 200000, dtype=np.uint8)),
             np.sort(rng.randint(0, 200000, 300000)).astype(np.int32).tobytes(),                  # a row member
             np.repeat(np.arange(2000, dtype=np.int32), rng.randint(1, 300, 2000)).tobytes(),     # a col member
             np.ones(200000, dtype=np.int64).tobytes(),                                           # a data member
             b"ab" * 70000 + b"xyz" * 50000 + b"q" * 100000 + b"0123456" * 30000 + b"01234" * 999 + b"012345" * 999,
             b" ".join(str(x).encode() for x in rng.randint(0, 10 ** 6, 100000))]
    for data in cases:
        for level in (0, 1, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
                for wbits in (-15, -9):
                    raw = _deflate(data, level, strategy, wbits)
                    out = su._native_inflate(raw, len(data))
                    assert out is not None and out.tobytes() == data and out.flags.writeable
                    if len(data) > 10:
                        assert su._native_inflate(raw, len(data) - 1) is None
                        assert su._native_inflate(raw, len(data) + 1) is None
                        assert su._native_inflate(raw[:len(raw) // 2], len(data)) is None
    # a stream with sync / full flushes (empty stored blocks between compressed ones)
    c, parts, data = zlib.compressobj(6, zlib.DEFLATED, -15), [], b""
    for k in range(40):
        chunk = bytes(rng.randint(0, 1 + k, rng.randint(1, 20000), dtype=np.uint8))
        data += chunk
        parts += [c.compress(chunk), c.flush(zlib.Z_SYNC_FLUSH if k % 2 else zlib.Z_FULL_FLUSH)]
    raw = b"".join(parts) + c.flush()
    assert su._native_inflate(raw, len(data)).tobytes() == data
    # damaged streams never crash; whatever comes back differs from the original or is refused
    big = np.sort(rng.randint(0, 200000, 100000)).astype(np.int32).tobytes()
    raw = _deflate(big)
    for _ in range(400):
        hurt = bytearray(raw)
        hurt[rng.randint(0, len(hurt))] ^= 1 << rng.randint(0, 8)
        out = su._native_inflate(bytes(hurt), len(big))
        assert out is None or zlib.crc32(out) != zlib.crc32(big) or out.tobytes() == big
    for _ in range(200):                                                                           # random garbage
        junk = bytes(rng.randint(0, 256, rng.randint(1, 400), dtype=np.uint8))
        su._native_inflate(junk, int(rng.randint(0, 5000)))


def test_read_lsdf_uses_the_native_decoder_and_falls_back(tmp_path, monkeypatch):
    coo = synth.bernoulli_matrix(3000, 40, 300, seed=5)
    labels = ["g%d" % i for i in range(3000)], ["s%d" % i for i in range(40)]
    path = str(tmp_path / "t.npz")
    su.LightSparseDataFrame(labels[0], labels[1], coo).to_npz(path)
    calls = []
    real = su._native_inflate
    monkeypatch.setattr(su, "_native_inflate", lambda raw, size: calls.append(size) or real(raw, size))
    fast = su.read_lsdf(path)
    assert len(calls) == 5 and (fast.data.tocsr() != coo.tocsr()).nnz == 0          # row, col, data, shape, format
    assert fast.data.row.flags.writeable and fast.data.data.dtype == np.int64
    monkeypatch.setattr(su, "_native_inflate", lambda raw, size: None)              # library absent / refusing
    slow = su.read_lsdf(path)
    assert (slow.data.tocsr() != coo.tocsr()).nnz == 0 and slow.data.row.flags.writeable
    wrong = lambda raw, size: np.zeros(size, dtype=np.uint8)                        # a wrong answer fails the CRC
    monkeypatch.setattr(su, "_native_inflate", wrong)
    assert (su.read_lsdf(path).data.tocsr() != coo.tocsr()).nnz == 0


def test_reference_import_paths_resolve_to_the_drop_in():
    """Both import styles of the reference (README.md:146 top-level modules inside the package directory,
    pangenome_analysis.py:22 ``import pangenomix.sparse_utils``) reach the B200 implementation."""
    import importlib
    import pangenomix_b200.core_genome
    import pangenomix_b200.pangenome_analysis
    import pangenomix_b200.plot
    import pangenomix_b200.sparse_utils
    for name in ("sparse_utils", "pangenome_analysis", "plot", "core_genome"):
        shim = importlib.import_module("pangenomix." + name)
        assert shim is getattr(pangenomix_b200, name)
    import pangenomix.pangenome_analysis as pa
    for fn in ("estimate_pan_core_size", "fit_heaps_by_iteration", "compute_bernoulli_grid_core_genome",
               "compute_beta_binomial_core_genome", "ks_montecarlo_bbn", "draw_bbn", "ecdf_from_counts", "betabin_logpmf"):
        assert callable(getattr(pa, fn)), fn
    import pangenomix.core_genome as cg
    assert callable(cg.count_gene_occurence)
