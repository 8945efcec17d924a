"""Host-side planning: COO presence/absence table -> the folded row layout of ``pgx_plan``.

Input is the object ``estimate_pan_core_size`` reads at
/root/reference/pangenomix/pangenome_analysis.py:74 (``df_genes.data``, a scipy COO
matrix, gene x genome).  The reference transposes it to genome-major CSR (:75) and walks
genomes; the CUDA path needs it gene-major, one row per gene, because a gene's
first-presence / first-absence rank is a reduction over that gene's genomes.

Everything here is numpy on the host, runs once per matrix ("uploaded once"), and is
O(nnz).  No compute of curves happens here.
"""
from __future__ import annotations

import dataclasses

import numpy as np
import scipy.sparse

MAX_GENOMES = 65535          # genome indices and ranks are uint16 on the device
CHUNK = 8                    # indices per 16-byte chunk
_COMPLEMENT_BLOCK_CELLS = 1 << 26


@dataclasses.dataclass
class HostPlan:
    n_genes: int
    n_genomes: int
    nnz: int
    chunks: np.ndarray        # uint16 [n_chunks * 8]
    row_ptr: np.ndarray       # int32  [n_rows + 1]   (chunk units)
    tasks: np.ndarray         # int32  [n_tasks, 2]
    w_present: np.ndarray     # int32  [N]
    w_absent: np.ndarray      # int32  [N]
    n_empty: int
    n_full: int
    row_gene: np.ndarray      # int64  [n_rows] gene id of every folded row (diagnostics / tests)
    row_len: np.ndarray       # int32  [n_rows] folded list length
    row_absent: np.ndarray    # bool   [n_rows] True when the list holds ABSENT genomes

    @property
    def n_rows(self):
        return int(self.row_len.shape[0])

    @property
    def n_tasks(self):
        return int(self.tasks.shape[0])

    @property
    def n_chunks(self):
        return int(self.chunks.shape[0] // CHUNK)

    @property
    def folded_nnz(self):
        return int(self.row_len.sum())

    @property
    def algorithmic_bytes_per_perm(self):
        """one pass over the canonical int32 gene-major CSR."""
        return 4 * self.nnz + 4 * (self.n_genes + 1)

    @property
    def streamed_bytes_per_pass(self):
        """Bytes the row kernel actually streams per pass over the rows (all B perms)."""
        return int(self.chunks.nbytes + self.row_ptr.nbytes + self.tasks.nbytes)


def gene_major_csr(data):
    """Canonical binary gene-major CSR of the table; raises on non-binary content.

    ``.tocsr()`` sums duplicate COO entries exactly as the reference's
    ``gene_data.T.tocsr()`` does (pangenome_analysis.py:75).  A summed value other than 1
    changes the reference's core curve but not its pan curve (SURVEY.md a-3'); no producer
    in the reference emits such tables (pangenome.py:631-650), so they are rejected
    rather than silently reinterpreted.
    """
    if isinstance(data, np.ndarray):
        data = scipy.sparse.coo_matrix(data)
    if not scipy.sparse.issparse(data):
        raise TypeError("expected a scipy sparse matrix or a dense ndarray, got %r" % type(data))
    csr = scipy.sparse.csr_matrix(data)
    csr.sum_duplicates()
    csr.sort_indices()
    if csr.nnz:
        values = np.asarray(csr.data)
        if not np.all(values == 1):
            if np.any(values == 0):
                csr.eliminate_zeros()
                values = np.asarray(csr.data)
            if not np.all(values == 1):
                raise ValueError(
                    "presence/absence table is not binary after summing duplicate entries "
                    "(found values other than 0/1); estimate_pan_core_size on the GPU "
                    "requires a binary table")
    return csr


def _segment_positions(lengths):
    """For concatenated segments of the given lengths: position of every element in its segment."""
    total = int(lengths.sum())
    starts = np.cumsum(lengths) - lengths
    return np.arange(total, dtype=np.int64) - np.repeat(starts, lengths)


def build_host_plan(data) -> HostPlan:
    csr = gene_major_csr(data)
    n_genes, n = csr.shape
    if n < 1:
        raise ValueError("table has no genome columns")
    if n > MAX_GENOMES:
        raise ValueError("n_genomes = %d exceeds the supported maximum of %d" % (n, MAX_GENOMES))
    if n_genes >= 2 ** 31 - 1:
        raise ValueError("too many genes")
    indptr = csr.indptr.astype(np.int64)
    indices = csr.indices
    m = np.diff(indptr)

    empty = m == 0
    full = (m == n) & ~empty
    single_p = (m == 1) & ~full
    single_a = (m == n - 1) & ~single_p & ~full & ~empty
    general = ~(empty | full | single_p | single_a)

    w_present = np.bincount(indices[indptr[:-1][single_p]], minlength=n).astype(np.int32)
    if single_a.any():
        csum = np.concatenate(([0], np.cumsum(indices, dtype=np.int64)))
        row_sum = csum[indptr[1:][single_a]] - csum[indptr[:-1][single_a]]
        missing = n * (n - 1) // 2 - row_sum
        w_absent = np.bincount(missing, minlength=n).astype(np.int32)
    else:
        w_absent = np.zeros(n, dtype=np.int32)

    genes = np.flatnonzero(general)
    m_gen = m[genes]
    use_abs = m_gen > n - m_gen
    length = np.where(use_abs, n - m_gen, m_gen).astype(np.int64)
    n_chunk = (length + CHUNK - 1) // CHUNK
    log_w = np.zeros(genes.shape[0], dtype=np.int64)
    for k in range(1, 6):
        log_w[n_chunk > (1 << (k - 1))] = k
    order = np.lexsort((-length, use_abs, -log_w))
    genes, use_abs, length, n_chunk, log_w = (a[order] for a in (genes, use_abs, length, n_chunk, log_w))

    row_ptr = np.concatenate(([0], np.cumsum(n_chunk)))
    if row_ptr[-1] >= 2 ** 31:
        raise ValueError("folded table too large for int32 chunk offsets")
    chunks = np.full(int(row_ptr[-1]) * CHUNK, n, dtype=np.uint16)
    dest_start = row_ptr[:-1] * CHUNK

    # rows that keep their present list: gather segments of csr.indices in the new order
    keep = np.flatnonzero(~use_abs)
    if keep.size:
        lens = length[keep]
        pos = _segment_positions(lens)
        src = np.repeat(indptr[genes[keep]], lens) + pos
        dst = np.repeat(dest_start[keep], lens) + pos
        chunks[dst] = indices[src].astype(np.uint16)
    # rows that store the complement: densify them block-wise
    comp = np.flatnonzero(use_abs)
    if comp.size:
        block = max(1, _COMPLEMENT_BLOCK_CELLS // n)
        for b0 in range(0, comp.size, block):
            rows = comp[b0:b0 + block]
            present_lens = m[genes[rows]]
            pos = _segment_positions(present_lens)
            src = np.repeat(indptr[genes[rows]], present_lens) + pos
            absent = np.ones((rows.size, n), dtype=bool)
            absent[np.repeat(np.arange(rows.size), present_lens), indices[src]] = False
            r_loc, cols = np.nonzero(absent)          # row-major => sorted inside each row
            lens = length[rows]
            assert r_loc.size == int(lens.sum())
            dst = np.repeat(dest_start[rows], lens) + _segment_positions(lens)
            chunks[dst] = cols.astype(np.uint16)

    # tasks: runs of rows sharing (lanes-per-row, list kind), 32 / lanes rows per warp task
    task_rows, task_meta = [], []
    if genes.size:
        key = log_w * 2 + use_abs
        change = np.flatnonzero(np.diff(key)) + 1
        run_starts = np.concatenate(([0], change))
        run_ends = np.concatenate((change, [genes.size]))
        for r0, r1 in zip(run_starts, run_ends):
            lw = int(log_w[r0])
            flag = int(use_abs[r0])
            per_task = 32 >> lw
            first = np.arange(r0, r1, per_task, dtype=np.int64)
            count = np.minimum(per_task, r1 - first)
            task_rows.append(first)
            task_meta.append((count << 8) | (lw << 1) | flag)
    if task_rows:
        tasks = np.stack([np.concatenate(task_rows), np.concatenate(task_meta)], axis=1).astype(np.int32)
    else:
        tasks = np.zeros((0, 2), dtype=np.int32)

    return HostPlan(
        n_genes=int(n_genes), n_genomes=int(n), nnz=int(csr.nnz),
        chunks=chunks, row_ptr=row_ptr.astype(np.int32), tasks=np.ascontiguousarray(tasks),
        w_present=w_present, w_absent=w_absent,
        n_empty=int(empty.sum()), n_full=int(full.sum()),
        row_gene=genes, row_len=length.astype(np.int32), row_absent=use_abs.astype(bool))
