"""Host-side planning: COO presence/absence table -> the device layout of ``pgx_plan``.

Input is the object ``estimate_pan_core_size`` reads at
/root/reference/pangenomix/pangenome_analysis.py:74 (``df_genes.data``, a scipy COO
matrix, gene x genome).  The reference transposes it to genome-major CSR (:75) and walks
genomes; the CUDA path needs it gene-major, one row per gene, because a gene's
first-presence / first-absence rank is a reduction over that gene's genomes.

Three populations of genes (include/pgx.h describes the device side):

* closed forms   -- empty / universal / single-genome / single-absence genes are functions of
                    perm[0] and of one rank; they become per-genome weight vectors.
* list rows      -- genes whose shorter list (present or absent genomes) has fewer than
                    ``long_threshold`` entries: uint16 genome indices in 16-byte chunks,
                    one lane per row, sub-blocks of 32 rows of equal chunk count, chunks of a
                    sub-block interleaved so that a warp reads 512 contiguous bytes per step
                    (a warp task streams a run of consecutive sub-blocks), and
                    the entries of every row ordered so that the lanes of a shared-memory
                    bank group hit distinct banks of the rank table.
* bitmap rows    -- the remaining (long) genes as a genome-major bit-sliced bitmap (32 genes per
                    word, 1,024 per 128-byte line); the probe kernel walks genomes in rank order
                    for 1,024 genes at once instead of scanning long lists.

Everything here is numpy on the host, runs once per matrix ("uploaded once"), and is
O(nnz).  No compute of curves happens here.
"""
from __future__ import annotations

import dataclasses

import numpy as np
import scipy.sparse

MAX_GENOMES = 65503          # genome indices, ranks and the 32 sentinel indices N..N+31 are uint16
CHUNK = 8                    # indices per 16-byte chunk
SENTINELS = 32               # rank-table rows N..N+31 hold 0xffff
SUPERBLOCK = 1024            # bitmap rows per (superblock, genome) line of 128 bytes
RUN_LANE_CHUNKS = 32         # most chunk iterations a warp streams per task (sub-blocks of 32 rows x chunks per row)
RUN_TARGET_TASKS = 1024      # ... but small tables keep enough tasks to spread over the warps of a CTA row
SMEM_TABLE_BUDGET = 220 * 1024


def default_long_threshold(n_genomes):
    """Folded length from which a gene is cheaper to serve from the bit-sliced bitmap (a walk of
    ~N/m coalesced lines shared by 1,024 genes) than from its list (m shared-memory gathers):
    the measured break-even on B200 is close to 1.3 sqrt(N) (profiles/)."""
    return max(8, int(round(1.28 * np.sqrt(n_genomes))))

_COMPLEMENT_BLOCK_CELLS = 1 << 26


def perms_per_cta_for(n_genomes):
    """Permutations whose rank tables fit one CTA's shared memory together (8, 4, 2 or 1)."""
    for b in (8, 4, 2):
        if (n_genomes + SENTINELS) * 2 * b <= SMEM_TABLE_BUDGET:
            return b
    return 1


def residue_modulus_for(perms_per_cta):
    """Lanes that share one 128-byte shared-memory wavefront when each gathers 2*B bytes."""
    return {8: 8, 4: 16, 2: 32, 1: 32}[int(perms_per_cta)]


@dataclasses.dataclass
class HostPlan:
    n_genes: int
    n_genomes: int
    nnz: int
    perms_per_cta: int
    long_threshold: int
    colsum: np.ndarray        # int32  [N]   genes present in genome c (all genes)
    w_present: np.ndarray     # int32  [N]   genes present ONLY in genome c
    w_absent: np.ndarray      # int32  [N]   genes absent ONLY from genome c
    n_empty: int
    n_full: int
    # list rows
    chunks: np.ndarray        # uint16 [n_chunks * 8]  task-interleaved, bank-ordered
    tasks: np.ndarray         # int32  [n_tasks, 4]  {first_chunk, chunks_per_row | absent<<24, first_row, rows}: a run of
                              #        ceil(rows / 32) consecutive sub-blocks of 32 rows x chunks_per_row chunks
    sorted_idx: np.ndarray    # uint16 [sum(row_len)]  the same lists, plainly sorted (mex probe)
    sorted_ptr: np.ndarray    # int32  [n_rows + 1]
    row_gene: np.ndarray      # int64  [n_rows] gene id of every list row (diagnostics / tests)
    row_len: np.ndarray       # int32  [n_rows] folded list length
    row_absent: np.ndarray    # bool   [n_rows] True when the list holds ABSENT genomes
    # bitmap rows
    bits: np.ndarray          # uint32 [n_superblocks * N * 32]  bit b of word (sb, c, w) = row 1024 sb + 32 w + b in genome c
    long_gene: np.ndarray     # int64  [n_long]
    nnz_list: int = 0         # present entries of the list-row genes (roofline apportioning)
    nnz_long: int = 0         # present entries of the bitmap-row genes

    @property
    def n_rows(self):
        return int(self.row_len.shape[0])

    @property
    def n_long(self):
        return int(self.long_gene.shape[0])

    @property
    def n_superblocks(self):
        return (self.n_long + SUPERBLOCK - 1) // SUPERBLOCK

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

    def algorithmic_bytes_of(self, kind):
        """The share of algorithmic_bytes_per_perm that belongs to the genes one row kernel serves."""
        if kind == "list":
            return 4 * self.nnz_list + 4 * self.n_rows
        if kind == "probe":
            return 4 * self.nnz_long + 4 * self.n_long
        raise KeyError(kind)

    @property
    def streamed_bytes_per_pass(self):
        """Bytes the row kernels stream per pass over the rows (list chunks + bitmaps)."""
        return int(self.chunks.nbytes + self.tasks.nbytes + self.bits.nbytes)


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


def _folded_lists(indptr, indices, m, genes, use_abs, length, n):
    """Concatenated sorted folded lists of ``genes`` (present or absent genomes), int32."""
    ptr = np.concatenate(([0], np.cumsum(length))).astype(np.int64)
    flat = np.empty(int(ptr[-1]), dtype=np.int32)
    keep = np.flatnonzero(~use_abs)
    if keep.size:
        lens = length[keep]
        pos = _segment_positions(lens)
        src = np.repeat(indptr[genes[keep]], lens) + pos
        dst = np.repeat(ptr[:-1][keep], lens) + pos
        flat[dst] = indices[src]
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
            _, cols = np.nonzero(absent)              # row-major => sorted inside each row
            lens = length[rows]
            assert cols.size == int(lens.sum())
            dst = np.repeat(ptr[:-1][rows], lens) + _segment_positions(lens)
            flat[dst] = cols
    return flat, ptr


def _bank_ordered_chunks(flat, ptr, task_first, task_nch, task_first_row, task_rows, n, modulus):
    """Lays the list rows out for the lane-per-row kernel.

    Lane l of a task reads chunk ``first + it * 32 + l`` at iteration ``it`` and gathers the
    rank-table line of the chunk's j-th index at step s = 8 it + j.  The lanes of one
    shared-memory wavefront (``modulus`` consecutive lanes) are conflict-free when their
    indices differ modulo ``modulus``, so an entry of residue r = index % modulus is stored
    at a step with (s + l) % modulus == r while the row has such steps left; surplus
    entries take the row's unused steps, and steps still unused keep a sentinel index
    (>= N) of the matching residue.
    """
    n_tasks = task_nch.shape[0]
    slots_per_task = task_nch * (32 * CHUNK)
    n_slots = int(slots_per_task.sum())
    slot = np.arange(n_slots, dtype=np.int64)
    task_of_slot = np.repeat(np.arange(n_tasks), slots_per_task)
    lane_of_slot = (slot >> 3) & 31                     # tasks start at multiples of 32 chunks
    it_of_slot = ((slot >> 3) - task_first[task_of_slot]) >> 5
    want = (it_of_slot * 8 + (slot & 7) + lane_of_slot) % modulus
    chunks = (n + ((want - n) % modulus)).astype(np.uint16)
    del want, it_of_slot
    if flat.size == 0:
        return chunks

    lens = np.diff(ptr)
    n_rows = lens.shape[0]
    row_task = np.repeat(np.arange(n_tasks), task_rows)
    row_lane = np.arange(n_rows) - task_first_row[row_task]
    row_of_entry = np.repeat(np.arange(n_rows), lens)
    res = flat % modulus
    order = np.lexsort((flat, res, row_of_entry))           # by (row, residue, index)
    row_s, res_s, idx_s = row_of_entry[order], res[order], flat[order]
    key = row_s * modulus + res_s
    starts = np.flatnonzero(np.concatenate(([True], key[1:] != key[:-1])))
    run_len = np.diff(np.concatenate((starts, [key.shape[0]])))
    occ = np.arange(key.shape[0]) - np.repeat(starts, run_len)
    lane = row_lane[row_s]
    step = ((res_s - lane) % modulus) + modulus * occ
    placed = step < task_nch[row_task[row_s]] * 8
    addr = ((task_first[row_task[row_s]] + (step >> 3) * 32 + lane) << 3) + (step & 7)
    filled = np.zeros(n_slots, dtype=bool)
    chunks[addr[placed]] = idx_s[placed].astype(np.uint16)
    filled[addr[placed]] = True

    over = np.flatnonzero(~placed)
    if over.size:
        # the k-th surplus entry of a row takes the row's k-th free slot
        over_row = row_s[over]                               # sorted by row already
        real = lane_of_slot < task_rows[task_of_slot]
        free = np.flatnonzero(real & ~filled)
        free_row = task_first_row[task_of_slot[free]] + lane_of_slot[free]
        by_row = np.argsort(free_row, kind="stable")
        free, free_row = free[by_row], free_row[by_row]
        free_start = np.searchsorted(free_row, np.arange(n_rows), side="left")
        over_start = np.searchsorted(over_row, np.arange(n_rows), side="left")
        k = np.arange(over.size) - over_start[over_row]
        dest = free[free_start[over_row] + k]
        assert np.array_equal(free_row[free_start[over_row] + k], over_row)
        chunks[dest] = idx_s[over].astype(np.uint16)
    return chunks


def build_host_plan(data, long_threshold=None, perms_per_cta=None) -> HostPlan:
    csr = gene_major_csr(data)
    n_genes, n = csr.shape
    if n < 1:
        raise ValueError("table has no genome columns")
    if n > MAX_GENOMES:
        raise ValueError("n_genomes = %d exceeds the supported maximum of %d" % (n, MAX_GENOMES))
    if n_genes >= 2 ** 31 - 1:
        raise ValueError("too many genes")
    if perms_per_cta is None:
        perms_per_cta = perms_per_cta_for(n)
    if perms_per_cta not in (1, 2, 4, 8):
        raise ValueError("perms_per_cta must be 1, 2, 4 or 8")
    modulus = residue_modulus_for(perms_per_cta)
    if long_threshold is None:
        long_threshold = default_long_threshold(n)
    long_threshold = int(long_threshold)
    indptr = csr.indptr.astype(np.int64)
    indices = csr.indices
    m = np.diff(indptr)
    colsum = np.bincount(indices, minlength=n).astype(np.int32)

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
    folded = np.minimum(m_gen, n - m_gen)
    if long_threshold > 0:
        is_long = folded >= long_threshold
    else:
        is_long = np.zeros(genes.shape[0], dtype=bool)

    # ---- bitmap rows: genome-major, bit-sliced, superblocks of 1,024 rows of similar density ----
    long_gene = genes[is_long]
    n_super = (long_gene.size + SUPERBLOCK - 1) // SUPERBLOCK
    if long_gene.size:
        m_long = m[long_gene]
        order = np.argsort(np.minimum(m_long, n - m_long), kind="stable")   # longest walks first
        long_gene = long_gene[order]
        bits = np.zeros((n_super, n, SUPERBLOCK // 32), dtype=np.uint32)
        for sb in range(n_super):
            rows = long_gene[sb * SUPERBLOCK:(sb + 1) * SUPERBLOCK]
            lens = m[rows]
            src = np.repeat(indptr[rows], lens) + _segment_positions(lens)
            dense = np.zeros((n, SUPERBLOCK), dtype=bool)                    # [genome][row in superblock]
            dense[indices[src], np.repeat(np.arange(rows.size), lens)] = True
            bits[sb] = np.packbits(dense, axis=1, bitorder="little").view(np.uint32)
        bits = bits.reshape(-1)
    else:
        bits = np.zeros(0, dtype=np.uint32)

    # ---- list rows -----------------------------------------------------------------------
    genes = genes[~is_long]
    m_gen = m[genes]
    use_abs = m_gen > n - m_gen
    length = np.where(use_abs, n - m_gen, m_gen).astype(np.int64)
    n_chunk = (length + CHUNK - 1) // CHUNK
    order = np.lexsort((-length, use_abs, -n_chunk))         # by chunk count (desc), kind, length
    genes, use_abs, length, n_chunk = (a[order] for a in (genes, use_abs, length, n_chunk))
    flat, ptr = _folded_lists(indptr, indices, m, genes, use_abs, length, n)
    if ptr[-1] >= 2 ** 31:
        raise ValueError("list rows too large for int32 offsets")

    n_rows = genes.shape[0]
    if n_rows:
        key = n_chunk * 2 + use_abs
        change = np.flatnonzero(np.diff(key)) + 1
        run_starts = np.concatenate(([0], change))
        run_ends = np.concatenate((change, [n_rows]))
        # sub-blocks of 32 rows (one lane per row) ...
        b_first_row, b_rows, b_nch, b_abs, b_run = [], [], [], [], []
        run_id = 0
        total_iters = int(((run_ends - run_starts + 31) // 32 * n_chunk[run_starts]).sum())
        run_budget = int(min(RUN_LANE_CHUNKS, max(2, total_iters // RUN_TARGET_TASKS)))
        for r0, r1 in zip(run_starts, run_ends):
            first_row = np.arange(r0, r1, 32, dtype=np.int64)
            nch = int(n_chunk[r0])
            b_first_row.append(first_row)
            b_rows.append(np.minimum(32, r1 - first_row))
            b_nch.append(np.full(first_row.shape[0], nch, dtype=np.int64))
            b_abs.append(np.full(first_row.shape[0], int(use_abs[r0]), dtype=np.int64))
            # ... streamed by a warp in runs of about RUN_LANE_CHUNKS chunk iterations
            per_run = max(1, run_budget // nch)
            b_run.append(run_id + np.arange(first_row.shape[0]) // per_run)
            run_id = int(b_run[-1][-1]) + 1
        b_first_row, b_rows, b_nch, b_abs, b_run = (np.concatenate(a) for a in (b_first_row, b_rows, b_nch, b_abs, b_run))
        block_first = np.concatenate(([0], np.cumsum(b_nch * 32)))[:-1]
        if (block_first[-1] + b_nch[-1] * 32) * CHUNK >= 2 ** 31:
            raise ValueError("list rows too large for int32 chunk offsets")
        if b_nch.max() >= 1 << 16:
            raise ValueError("list row too long for the task descriptor")
        chunks = _bank_ordered_chunks(flat, ptr, block_first, b_nch, b_first_row, b_rows, n, modulus)
        head = np.flatnonzero(np.concatenate(([True], np.diff(b_run) != 0)))       # first sub-block of every run
        rows_in_run = np.add.reduceat(b_rows, head)
        tasks = np.stack([block_first[head], b_nch[head] | (b_abs[head] << 24), b_first_row[head], rows_in_run],
                         axis=1).astype(np.int32)
        # costly runs first: dynamic fetching then ends on cheap ones
        cost = b_nch[head] * ((rows_in_run + 31) // 32)
        tasks = tasks[np.argsort(-cost, kind="stable")]
    else:
        chunks = np.zeros(0, dtype=np.uint16)
        tasks = np.zeros((0, 4), dtype=np.int32)

    return HostPlan(
        n_genes=int(n_genes), n_genomes=int(n), nnz=int(csr.nnz),
        perms_per_cta=int(perms_per_cta), long_threshold=long_threshold,
        colsum=colsum, w_present=w_present, w_absent=w_absent,
        n_empty=int(empty.sum()), n_full=int(full.sum()),
        chunks=chunks, tasks=np.ascontiguousarray(tasks),
        sorted_idx=flat.astype(np.uint16), sorted_ptr=ptr.astype(np.int32),
        row_gene=genes, row_len=length.astype(np.int32), row_absent=use_abs.astype(bool),
        bits=bits, long_gene=long_gene,
        nnz_list=int(m[genes].sum()), nnz_long=int(m[long_gene].sum()))
