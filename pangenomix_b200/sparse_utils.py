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


def read_lsdf(npz_file, label_file=None):
    """Loads an LSDF written by ``LightSparseDataFrame.to_npz`` (sparse_utils.py:18-42).

    ``label_file`` defaults to ``<npz_file>.labels.txt``: one label per line, the row
    labels first, then the column labels.
    """
    matrix = scipy.sparse.load_npz(npz_file)
    if label_file is None:
        label_file = npz_file + ".labels.txt"
    with open(label_file, "r") as handle:
        labels = [line.strip() for line in handle]
    n_rows = matrix.shape[0]
    return LightSparseDataFrame(labels[:n_rows], labels[n_rows:], matrix)


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
