"""Drop-in for the one marginal of /root/reference/pangenomix/core_genome.py that lies on the table's data-parallel
path: ``count_gene_occurence`` (:127-155).  The other functions of that module work
on FASTA / label files and are out of scope."""
from __future__ import print_function

import numpy as np
import pandas as pd

from .sparse_utils import _load_npz_coo


def count_gene_occurence(gene_npz_file, device=None):
    '''
    Number of genomes every gene of a ``<name>_strain_by_gene.npz`` table occurs in: a DataFrame with columns
    gene_index (the row index of the gene, dtype of the stored ``row`` array) and count (int64), genes that
    occur nowhere left out, ascending gene index -- as the reference's groupby / drop_duplicates / sort builds it
    (:131-151).  Counted on the GPU (pgx_coo_marginals_host).
    '''
    from .engine import table_marginals
    coo = _load_npz_coo(gene_npz_file)
    if coo is not None:
        row, col, shape = coo.row, coo.col, coo.shape
    else:                                            # not an archive to_npz wrote: read it as the reference does
        with np.load(gene_npz_file) as data:
            row, col = data['row'], data['col']
        shape = (int(row.max()) + 1 if row.size else 0, int(col.max()) + 1 if col.size else 0)
    row_sum, _, _, _ = table_marginals(_Entries(row, col, shape), device=device, spectrum=False)
    genes = np.flatnonzero(row_sum)
    df_unique = pd.DataFrame({'gene_index': genes.astype(row.dtype), 'count': row_sum[genes]})
    print("\nCounted gene occurence")
    return df_unique


class _Entries:
    """The reference counts stored ENTRIES per row index whatever their values (:137): present the table so."""

    def __init__(self, row, col, shape):
        self._row, self._col, self.shape = row, col, tuple(shape)

    def tocoo(self):
        import scipy.sparse
        ones = np.ones(self._row.shape[0], dtype=np.int8)
        return scipy.sparse.coo_matrix((ones, (self._row, self._col)), shape=self.shape)
