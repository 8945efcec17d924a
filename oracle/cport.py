"""TEST INFRASTRUCTURE -- ctypes view of oracle/_build/liboracle.so (oracle/pancore_ref.c)."""
import ctypes
import os

import numpy as np
import scipy.sparse

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.OUT
        if not os.path.exists(path):
            path = _build.build()
        _lib = ctypes.CDLL(path)
        _lib.pgx_oracle_curves_direct.restype = ctypes.c_int
        _lib.pgx_oracle_curves_direct.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _lib.pgx_oracle_legacy_shuffles.restype = None
        _lib.pgx_oracle_legacy_shuffles.argtypes = [
            ctypes.c_uint32, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]
        _lib.pgx_oracle_max_threads.restype = ctypes.c_int
    return _lib


class GenomeMajor:
    """``df_genes.data.T.tocsr()`` (pangenome_analysis.py:74-75) held as C-ready arrays."""

    def __init__(self, data):
        csr = scipy.sparse.coo_matrix(data).T.tocsr()
        self.n_genomes, self.n_genes = csr.shape
        self.indptr = np.ascontiguousarray(csr.indptr, dtype=np.int32)
        self.indices = np.ascontiguousarray(csr.indices, dtype=np.int32)
        self.values = np.ascontiguousarray(csr.data, dtype=np.int64)


def curves_direct(data, perms, n_threads=1):
    """(pan, core) float64 via the C port of pangenome_analysis.py:81-90."""
    gm = data if isinstance(data, GenomeMajor) else GenomeMajor(data)
    perms = np.ascontiguousarray(perms, dtype=np.int32)
    n_iter = perms.shape[0]
    pan = np.zeros((n_iter, gm.n_genomes))
    core = np.zeros((n_iter, gm.n_genomes))
    rc = lib().pgx_oracle_curves_direct(
        gm.indptr.ctypes.data, gm.indices.ctypes.data, gm.values.ctypes.data,
        gm.n_genomes, gm.n_genes, perms.ctypes.data, n_iter,
        pan.ctypes.data, core.ctypes.data, int(n_threads))
    if rc != 0:
        raise MemoryError("oracle C port could not allocate its incidence vector")
    return pan, core


def legacy_shuffles(seed, n, count):
    out = np.empty((count, n), dtype=np.int32)
    lib().pgx_oracle_legacy_shuffles(int(seed), int(n), int(count), out.ctypes.data)
    return out


def max_threads():
    return int(lib().pgx_oracle_max_threads())
