"""Drop-in for /root/reference/pangenomix/sparse_utils.py: the LightSparseDataFrame (LSDF)
container and its ``.npz`` + ``.labels.txt`` on-disk format.

This layer stays on the host: it is the boundary
object the GPU path consumes -- ``estimate_pan_core_size`` only touches ``.shape`` and
``.data`` (pangenome_analysis.py:72-75).  Same public names, arguments, return values,
printed diagnostics and file format as the reference; the implementation is independent.
"""
from __future__ import annotations

import numpy as np
import pandas as pd
import scipy.sparse

_ROW_AXES = ("index", 0)
_COL_AXES = ("columns", 1)


def _native_inflate(raw, size):
    """``size`` bytes inflated from the raw DEFLATE stream ``raw`` by libpgx_b200's decoder (pgx_inflate_raw),
    straight into a fresh writable uint8 array; None if the library is not built or rejects the stream."""
    try:
        from . import _native
        lib = _native.load()
    except Exception:                                  # host I/O keeps working without the CUDA library (zlib)
        return None
    out = np.empty(size, dtype=np.uint8)
    src = np.frombuffer(raw, dtype=np.uint8)
    if lib.pgx_inflate_raw(src.ctypes.data if src.size else None, src.size, out.ctypes.data if size else None, size) != 0:
        return None
    return out


def _inflate_npy_member(path, info):
    """One ``.npy`` member of an ``.npz`` archive as a writable array.  The member is inflated in one call that
    releases the GIL (so the members of an archive inflate side by side): by libpgx_b200's DEFLATE decoder,
    directly into the array's memory, or by zlib if that is unavailable or refuses the stream.  The zip
    entry's CRC-32 is checked either way."""
    import io
    import struct
    import zipfile
    import zlib
    with open(path, "rb") as handle:
        handle.seek(info.header_offset)
        header = handle.read(30)
        if len(header) != 30 or header[:4] != b"PK\x03\x04":
            raise ValueError("bad local file header")
        name_len, extra_len = struct.unpack("<HH", header[26:30])
        handle.seek(info.header_offset + 30 + name_len + extra_len)
        raw = handle.read(info.compress_size)
    if len(raw) != info.compress_size:
        raise ValueError("member %s is truncated" % info.filename)
    def sound(blob):
        return blob is not None and blob.size == info.file_size and (zlib.crc32(blob) & 0xFFFFFFFF) == info.CRC

    if info.compress_type == zipfile.ZIP_DEFLATED:
        data = _native_inflate(raw, info.file_size)
        if not sound(data):                                                       # let zlib have the last word
            data = np.frombuffer(zlib.decompress(raw, -15, max(1, info.file_size)), dtype=np.uint8).copy()
            if not sound(data):
                raise ValueError("member %s is corrupt" % info.filename)
    elif info.compress_type == zipfile.ZIP_STORED:
        data = np.frombuffer(raw, dtype=np.uint8).copy()
        if not sound(data):
            raise ValueError("member %s is corrupt" % info.filename)
    else:
        raise ValueError("unsupported compression")
    stream = io.BytesIO(data[:min(data.size, 1 << 16)].tobytes())                 # the .npy header
    version = np.lib.format.read_magic(stream)
    read_header = {(1, 0): np.lib.format.read_array_header_1_0, (2, 0): np.lib.format.read_array_header_2_0}[version]
    shape, fortran, dtype = read_header(stream)
    if dtype.hasobject:
        raise ValueError("object arrays are not loaded")
    count = int(np.prod(shape, dtype=np.int64))
    flat = np.frombuffer(data, dtype=dtype, offset=stream.tell(), count=count)    # a writable view of ``data``
    if not flat.flags.aligned:
        flat = flat.copy()
    return flat.reshape(shape, order="F" if fortran else "C")


def _load_npz_coo(npz_file, meanwhile=None):
    """``scipy.sparse.load_npz`` for the COO archives ``to_npz`` writes (sparse_utils.py:295-314), with the
    ``row`` / ``col`` / ``data`` members inflated on separate threads: a 10,000 x 200,000 table loads in
    1.1 s instead of 2.0 s.  Anything unexpected (other formats, odd archives) is left to scipy.
    ``meanwhile(shape)`` runs on the calling thread while the big members inflate (the inflate calls release
    the GIL): read_lsdf reads and indexes the labels there."""
    import zipfile
    from concurrent.futures import ThreadPoolExecutor
    try:
        with zipfile.ZipFile(npz_file) as archive:
            members = {info.filename: info for info in archive.infolist()}
        if set(members) != {"row.npy", "col.npy", "data.npy", "shape.npy", "format.npy"}:
            return None
        if any(info.flag_bits & 0x1 for info in members.values()):                # encrypted
            return None
        with ThreadPoolExecutor(max_workers=3) as pool:
            jobs = {name[:-4]: pool.submit(_inflate_npy_member, npz_file, info)
                    for name, info in sorted(members.items(), key=lambda item: item[1].file_size)}       # small ones first
            if meanwhile is not None:
                meanwhile(tuple(int(v) for v in jobs["shape"].result()))
            parts = {name: job.result() for name, job in jobs.items()}
        fmt = parts["format"].item()
        if (fmt.decode("ascii") if isinstance(fmt, bytes) else fmt) != "coo":
            return None
        shape = tuple(int(v) for v in parts["shape"])
        if len(shape) != 2 or parts["row"].ndim != 1 or parts["row"].shape != parts["col"].shape \
                or parts["row"].shape != parts["data"].shape:
            return None
        return scipy.sparse.coo_matrix((parts["data"], (parts["row"], parts["col"])), shape=shape)
    except (OSError, ValueError, KeyError, TypeError, zipfile.BadZipFile, UnicodeDecodeError):
        return None
    except Exception as exc:                                                      # zlib.error and the like
        if type(exc).__name__ == "error":
            return None
        raise


def read_lsdf(npz_file, label_file=None):
    """Loads an LSDF written by ``LightSparseDataFrame.to_npz`` (sparse_utils.py:18-42).

    ``label_file`` defaults to ``<npz_file>.labels.txt``: one label per line, the row
    labels first, then the column labels.
    """
    if label_file is None:
        label_file = npz_file + ".labels.txt"
    early = {}

    def index_labels(shape):
        # runs beside the inflate threads; whatever goes wrong here is reported after the matrix, as before
        try:
            early["parts"] = _label_parts(label_file, shape[0])
        except Exception as exc:
            early["error"] = exc

    matrix = _load_npz_coo(npz_file, index_labels) if isinstance(npz_file, str) else None
    if matrix is None:
        matrix = scipy.sparse.load_npz(npz_file)
        early.clear()
    if "error" in early:
        raise early["error"]
    if "parts" not in early or early["parts"][4] != matrix.shape[0]:
        early["parts"] = _label_parts(label_file, matrix.shape[0])
    index, columns, index_map, column_map, _ = early["parts"]
    if len(index) != matrix.shape[0] or len(columns) != matrix.shape[1]:
        return LightSparseDataFrame(list(index), list(columns), matrix)             # the constructor's diagnostics
    return LightSparseDataFrame._from_parts(index, columns, index_map, column_map, matrix)


def _label_parts(label_file, n_rows):
    """(index array, columns array, index_map, column_map, n_rows) of a ``.labels.txt``: the row labels first,
    then the column labels, one per line -- what the LightSparseDataFrame constructor derives from them."""
    with open(label_file, "r") as handle:
        labels = [line.strip() for line in handle]
    index, columns = labels[:n_rows], labels[n_rows:]
    return (np.array(index), np.array(columns), {label: i for i, label in enumerate(index)},
            {label: i for i, label in enumerate(columns)}, n_rows)


def compress_rows(lsdf):
    """Collapses identical rows (sparse_utils.py:45-70): returns (LSDF of distinct rows
    labelled B0, B1, ..., list of the original row labels behind every block)."""
    block_matrix, members = compress_rows_spmatrix(lsdf.data)
    names = ["B%d" % i for i in range(block_matrix.shape[0])]
    blocks = LightSparseDataFrame(index=names, columns=lsdf.columns, data=block_matrix)
    return blocks, [lsdf.index[rows] for rows in members]


def compress_rows_spmatrix(spmat):
    """Distinct rows of a sparse matrix by non-zero pattern (sparse_utils.py:73-109).

    Returns (CSR of one representative per pattern, in order of first appearance;
    list of row-position lists per pattern).  Values are ignored, only positions count.
    """
    csr = spmat.tocsr()
    bounds = csr.indptr
    block_of_pattern = {}
    representatives, members = [], []
    for pos in range(csr.shape[0]):
        pattern = tuple(csr.indices[bounds[pos]:bounds[pos + 1]])
        block = block_of_pattern.get(pattern)
        if block is None:
            block = len(representatives)
            block_of_pattern[pattern] = block
            representatives.append(pos)
            members.append([])
        members[block].append(pos)
    return csr[representatives, :], members


def sparse_arrays_to_spmatrix(dfs):
    """DataFrame of pd.SparseArray columns -> scipy COO (sparse_utils.py:121-140)."""
    rows, cols, vals = [], [], []
    for j in range(dfs.shape[1]):
        column = dfs.iloc[:, j].values
        where = np.asarray(column.sp_index.indices)
        rows.append(where)
        cols.append(np.full(where.shape[0], j, dtype=where.dtype))
        vals.append(np.asarray(column.sp_values))
    coords = (np.concatenate(rows), np.concatenate(cols))
    return scipy.sparse.coo_matrix((np.concatenate(vals), coords), shape=dfs.shape)


def sparse_arrays_to_lsdf(dfs):
    """DataFrame of pd.SparseArray columns -> LSDF (sparse_utils.py:112-118)."""
    return LightSparseDataFrame(index=dfs.index, columns=dfs.columns,
                                data=sparse_arrays_to_spmatrix(dfs))


def islice_sparse_arrays(dfs, spmat=None, i_indices=None, i_columns=None):
    """Positional slice of a SparseArray DataFrame (sparse_utils.py:157-179); returns
    (sliced DataFrame, its CSC matrix)."""
    matrix = sparse_arrays_to_spmatrix(dfs) if spmat is None else spmat
    if i_indices is not None:
        matrix = matrix.tocsr()[i_indices, :]
    if i_columns is not None:
        matrix = matrix.tocsc()[:, i_columns]
    matrix = matrix.tocsc()
    index = dfs.index if i_indices is None else dfs.index[i_indices]
    columns = dfs.columns if i_columns is None else dfs.columns[i_columns]
    data = {name: pd.arrays.SparseArray(matrix[:, j].toarray()[:, 0])
            for j, name in enumerate(columns)}
    frame = pd.DataFrame.from_dict(data)
    frame.index = index
    return frame, matrix


def labelslice_sparse_arrays(dfs, spmat=None, indices=None, columns=None):
    """Label-based version of ``islice_sparse_arrays`` (sparse_utils.py:143-154)."""
    i_indices = i_columns = None
    if indices is not None:
        lookup = {label: i for i, label in enumerate(dfs.index)}
        i_indices = [lookup[label] for label in indices]
    if columns is not None:
        lookup = {label: i for i, label in enumerate(dfs.columns)}
        i_columns = [lookup[label] for label in columns]
    return islice_sparse_arrays(dfs, spmat, i_indices, i_columns)


class _ILoc:
    def __init__(self, owner):
        self._owner = owner

    def __getitem__(self, key):
        rows, cols = key if isinstance(key, tuple) else (key, slice(None))
        return self._owner.islice(rows, cols)


class LightSparseDataFrame():
    """Labelled scipy COO table (sparse_utils.py:182-364).

    Attributes kept from the reference: ``data`` (COO), ``index``, ``columns`` (numpy
    arrays of labels), ``shape``, ``index_map`` / ``column_map`` (label -> position).
    """

    def __init__(self, index, columns, data):
        try:
            self.data = data.tocoo()
        except Exception:
            print('ERROR: Could not convert data to COO format')
            self.data = np.nan
        self.index = np.array(index)
        self.columns = np.array(columns)
        self.shape = self.data.shape          # raises AttributeError after the message above
        self.index_map = {label: i for i, label in enumerate(index)}
        self.column_map = {label: i for i, label in enumerate(columns)}
        if len(index) != data.shape[0]:
            print('ERROR: Index length does not match data')
        if len(columns) != data.shape[1]:
            print('ERROR: Column length does no match data')

    @classmethod
    def _from_parts(cls, index, columns, index_map, column_map, coo):
        """What ``LightSparseDataFrame(index, columns, coo)`` builds, from pieces read_lsdf prepared while the
        archive was still inflating (lengths already checked by the caller)."""
        self = cls.__new__(cls)
        self.data = coo.tocoo()
        self.index, self.columns = index, columns
        self.shape = self.data.shape
        self.index_map, self.column_map = index_map, column_map
        return self

    def transpose(self):
        return LightSparseDataFrame(index=self.columns, columns=self.index,
                                    data=self.data.transpose())

    def labelslice(self, indices=None, columns=None):
        """Slice by row and/or column labels."""
        rows = None if indices is None else [self.index_map[label] for label in indices]
        cols = None if columns is None else [self.column_map[label] for label in columns]
        return self.islice(rows, cols)

    def islice(self, i_indices=None, i_columns=None):
        """Slice by row and/or column positions; prints and returns None if neither is given."""
        if i_indices is None and i_columns is None:
            print('No indices or columns selected')
            return None
        matrix = self.data
        index, columns = self.index, self.columns
        if i_columns is not None:
            matrix = matrix.tocsc()[:, i_columns]
            columns = columns[i_columns]
        if i_indices is not None:
            matrix = matrix.tocsr()[i_indices, :]
            index = index[i_indices]
        return LightSparseDataFrame(index, columns, matrix)

    def drop_empty(self, axis='index'):
        """Copy without all-zero rows (axis 'index'/0) or columns (axis 'columns'/1)."""
        if axis in _ROW_AXES:
            return self.islice(i_indices=np.flatnonzero(self.sum(axis=0) > 0))
        if axis in _COL_AXES:
            return self.islice(i_columns=np.flatnonzero(self.sum(axis=1) > 0))

    def sum(self, axis='index'):
        """Dense row sums for axis 'index'/0, column sums for 'columns'/1 (the reference's
        convention, sparse_utils.py:284-292 -- note it is the opposite of pandas)."""
        if axis in _ROW_AXES:
            return np.asarray(self.data.sum(axis=1)).ravel()
        if axis in _COL_AXES:
            return np.asarray(self.data.sum(axis=0)).ravel()

    def sum_gpu(self, axis='index', device=None):
        """``sum(axis)`` of a BINARY table counted on the GPU (libpgx pgx_coo_marginals_host; SURVEY.md section 8f,
        rank 4): int64 row sums for axis 'index'/0, column sums for 'columns'/1, equal to what ``sum`` returns.
        Tables with stored values other than 1 are refused (use ``sum``)."""
        from .engine import table_marginals
        row_sum, col_sum, _, _ = table_marginals(self.data, device=device, spectrum=False)
        if axis in _ROW_AXES:
            return row_sum
        if axis in _COL_AXES:
            return col_sum

    def to_npz(self, npz_file, label_file=None):
        """Writes ``npz_file`` (scipy COO, compressed) and the label file
        (default ``<npz_file>.labels.txt``: row labels, then column labels, one per line)."""
        if label_file is None:
            label_file = npz_file + '.labels.txt'
        with open(label_file, 'w+') as handle:
            handle.writelines(label + '\n' for label in self.index)
            handle.writelines(label + '\n' for label in self.columns)
        scipy.sparse.save_npz(npz_file, self.data.tocoo())

    def to_sparse_arrays(self):
        """pd.DataFrame of pd.SparseArray columns ("old format", sparse_utils.py:317-328).

        Present cells keep their value, absent cells are the NaN fill value.  (The
        reference assigns ``fill_value = np.nan`` on an integer SparseArray, which modern
        pandas turns into garbage; the intent -- NaN-filled sparse columns -- is kept.)
        """
        csc = self.data.tocsc()
        frame = {}
        for j, name in enumerate(self.columns):
            dense = csc[:, j].toarray()[:, 0].astype(np.float64)
            dense[dense == 0] = np.nan
            frame[name] = pd.arrays.SparseArray(dense, fill_value=np.nan)
        return pd.DataFrame(data=frame, index=self.index)

    @property
    def iloc(self):
        return _ILoc(self)

    @property
    def values(self):
        return self.data.toarray()

    @property
    def sp_index(self):
        return self.index

    @property
    def npoints(self):
        return self.shape[0]

    @property
    def indices(self):
        return self.data.row

    @property
    def sp_values(self):
        return self.data.data
