// Host-side helper of the planner (pangenomix_b200/plan.py): the bank ordering of the list rows.
//
// plan._bank_ordered_chunks decides, for every list row (the folded genome list of one gene, see
// include/pgx.h), which entry sits at which gather step so that the lanes of a shared-memory
// wavefront of list_kernel hit distinct banks of the rank table.  The numpy version in plan.py is
// the specification (and stays, as the cross-check of the tests); this is the same algorithm --
// same greedy edge colouring, same tie breaks, bit-identical output -- as plain loops over the
// sub-blocks, spread over host threads.  It is O(slots) and removes three quarters of the planning
// time of large tables.  Nothing here touches the GPU.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "pgx.h"

namespace pgx {
int fail(int code, const char *fmt, ...);
}

namespace {

constexpr int MAX_MOD = 32;

struct Group {
    int r_mod, n_steps;
    int cnt[MAX_MOD][MAX_MOD];             // [lane][residue]
    std::vector<int8_t> res_at, pad_res;   // [lane][step]
};

// plan._colour_groups for one group: greedy edge colouring of lanes x residues with n_steps colours.
void colour(Group &g)
{
    const int R = g.r_mod, S = g.n_steps;
    int cnt[MAX_MOD][MAX_MOD], colload[MAX_MOD], rem[MAX_MOD];
    memcpy(cnt, g.cnt, sizeof(cnt));
    for (int r = 0; r < R; ++r) {
        colload[r] = 0;
        for (int l = 0; l < R; ++l) colload[r] += cnt[l][r];
    }
    for (int l = 0; l < R; ++l) {
        rem[l] = 0;
        for (int r = 0; r < R; ++r) rem[l] += cnt[l][r];
    }
    for (int s = 0; s < S; ++s) {
        bool taken[MAX_MOD] = {false};
        const int steps_left = S - s;
        // rows with the least slack choose first (stable argsort of steps_left - rem)
        int order[MAX_MOD];
        for (int l = 0; l < R; ++l) order[l] = l;
        std::stable_sort(order, order + R, [&](int a, int b) { return steps_left - rem[a] < steps_left - rem[b]; });
        for (int t = 0; t < R; ++t) {
            const int lane = order[t];
            int choice = 0;
            long best = -2;
            bool has = false;
            for (int r = 0; r < R; ++r) {              // argmax of colload * 1024 + cnt over available residues (first max wins)
                const bool avail = cnt[lane][r] > 0 && !taken[r];
                const long score = avail ? static_cast<long>(colload[r]) * 1024 + cnt[lane][r] : -1;
                if (score > best) {
                    best = score;
                    choice = r;
                    has = avail;
                }
            }
            const bool forced = !has && rem[lane] >= steps_left && rem[lane] > 0;
            if (forced) {
                int top = -1;
                for (int r = 0; r < R; ++r)
                    if (cnt[lane][r] > top) {
                        top = cnt[lane][r];
                        choice = r;
                    }
            }
            if (has || forced) {
                g.res_at[lane * S + s] = static_cast<int8_t>(choice);
                --cnt[lane][choice];
                --colload[choice];
                --rem[lane];
                taken[choice] = true;
            }
        }
        // pads: the k-th idle lane of the group takes the group's k-th unused residue
        int free_order[MAX_MOD], n_free = 0;
        for (int r = 0; r < R; ++r)
            if (!taken[r]) free_order[n_free++] = r;
        for (int r = 0; r < R; ++r)
            if (taken[r]) free_order[n_free++] = r;
        int rank = 0;
        for (int l = 0; l < R; ++l) {
            if (g.res_at[l * S + s] < 0) {
                g.pad_res[l * S + s] = static_cast<int8_t>(free_order[std::min(rank, R - 1)]);
                ++rank;
            }
        }
    }
}

// plan._positional_groups for one group.
void positional(Group &g)
{
    const int R = g.r_mod, S = g.n_steps;
    for (int l = 0; l < R; ++l) {
        int cap[MAX_MOD] = {0};
        for (int s = 0; s < S; ++s) {
            const int want = (s + l) % R;
            ++cap[want];
            g.pad_res[l * S + s] = static_cast<int8_t>(want);
            if (s / R < g.cnt[l][want]) g.res_at[l * S + s] = static_cast<int8_t>(want);
        }
        int s_free = 0;
        for (int r = 0; r < R; ++r) {
            for (int extra = std::max(g.cnt[l][r] - cap[r], 0); extra > 0; --extra) {
                while (s_free < S && g.res_at[l * S + s_free] >= 0) ++s_free;
                if (s_free < S) g.res_at[l * S + s_free] = static_cast<int8_t>(r);
            }
        }
    }
}

}  // namespace

extern "C" int pgx_plan_bank_order(const int32_t *flat, const int64_t *ptr, int64_t n_rows,
                                   const int64_t *block_first, const int64_t *block_nch,
                                   const int64_t *block_first_row, const int64_t *block_rows, int64_t n_blocks,
                                   int32_t n_genomes, int32_t modulus, int32_t colour_max_chunks,
                                   uint16_t *chunks, int32_t n_threads)
{
    if (!ptr || !block_first || !block_nch || !block_first_row || !block_rows || (!chunks && n_blocks > 0))
        return pgx::fail(PGX_ERR_INVALID, "null pointer passed to pgx_plan_bank_order");
    if (modulus != 8 && modulus != 16 && modulus != 32) return pgx::fail(PGX_ERR_INVALID, "modulus must be 8, 16 or 32");
    if (n_rows < 0 || n_blocks < 0) return pgx::fail(PGX_ERR_INVALID, "negative size");
    const int R = modulus;
    std::atomic<long long> next{0};
    std::atomic<int> bad{0};
    auto work = [&]() {
        Group g;
        g.r_mod = R;
        std::vector<int32_t> byres;                    // entries of one row, grouped by residue, index order kept
        for (;;) {
            const long long b = next.fetch_add(1);
            if (b >= n_blocks) return;
            const int S = static_cast<int>(block_nch[b]) * 8;
            g.n_steps = S;
            g.res_at.assign(static_cast<size_t>(R) * S, -1);
            g.pad_res.assign(static_cast<size_t>(R) * S, 0);
            for (int grp = 0; grp < 32 / R; ++grp) {
                std::fill(g.res_at.begin(), g.res_at.end(), static_cast<int8_t>(-1));
                std::fill(g.pad_res.begin(), g.pad_res.end(), static_cast<int8_t>(0));
                memset(g.cnt, 0, sizeof(g.cnt));
                for (int l = 0; l < R; ++l) {
                    const long long lane32 = grp * R + l;
                    if (lane32 >= block_rows[b]) continue;
                    const long long row = block_first_row[b] + lane32;
                    if (row >= n_rows || ptr[row + 1] - ptr[row] > S) {
                        bad.store(1);
                        continue;
                    }
                    for (long long e = ptr[row]; e < ptr[row + 1]; ++e) ++g.cnt[l][flat[e] % R];
                }
                if (block_nch[b] <= colour_max_chunks) colour(g); else positional(g);
                for (int l = 0; l < R; ++l) {
                    const long long lane32 = grp * R + l;
                    const bool real = lane32 < block_rows[b];
                    int start[MAX_MOD + 1];
                    start[0] = 0;
                    for (int r = 0; r < R; ++r) start[r + 1] = start[r] + g.cnt[l][r];
                    if (real) {
                        const long long row = block_first_row[b] + lane32;
                        byres.resize(static_cast<size_t>(start[R]));
                        int at[MAX_MOD];
                        for (int r = 0; r < R; ++r) at[r] = start[r];
                        for (long long e = ptr[row]; e < ptr[row + 1]; ++e) byres[at[flat[e] % R]++] = flat[e];
                    }
                    int used[MAX_MOD] = {0};
                    for (int s = 0; s < S; ++s) {
                        const long long addr = ((block_first[b] + static_cast<long long>(s >> 3) * 32 + lane32) << 3) + (s & 7);
                        const int r = g.res_at[l * S + s];
                        if (r >= 0 && real && used[r] < g.cnt[l][r]) {
                            chunks[addr] = static_cast<uint16_t>(byres[start[r] + used[r]++]);
                        } else {
                            const int pad = g.pad_res[l * S + s];
                            chunks[addr] = static_cast<uint16_t>(n_genomes + (((pad - n_genomes) % R) + R) % R);
                        }
                    }
                    for (int r = 0; r < R; ++r)
                        if (real && used[r] != g.cnt[l][r]) bad.store(1);
                }
            }
        }
    };
    int threads = n_threads > 0 ? n_threads : static_cast<int>(std::min(16u, std::max(1u, std::thread::hardware_concurrency())));
    threads = static_cast<int>(std::min<long long>(threads, std::max<long long>(1, n_blocks / 64)));
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work);
    work();
    for (auto &th : pool) th.join();
    if (bad.load()) return pgx::fail(PGX_ERR_INVALID, "inconsistent list rows passed to pgx_plan_bank_order");
    return PGX_OK;
}

// The bit-sliced bitmap of the long rows (struct pgx_plan, d_bits): bit b of word
// (sb * n_genomes + c) * 32 W + w is the presence of long row sb * 1024 W + 32 w + b in genome c.
// ``bits`` must be zero-initialised; superblocks are independent and spread over host threads.
extern "C" int pgx_plan_build_bitmap(const int64_t *indptr, const int32_t *indices, const int64_t *long_gene,
                                     int64_t n_long, int32_t n_genomes, int32_t slice_words, uint32_t *bits,
                                     int32_t n_threads)
{
    if (n_long < 0 || n_genomes < 1 || (slice_words != 1 && slice_words != 2 && slice_words != 4))
        return pgx::fail(PGX_ERR_INVALID, "bad shape passed to pgx_plan_build_bitmap");
    if (n_long == 0) return PGX_OK;
    if (!indptr || !indices || !long_gene || !bits) return pgx::fail(PGX_ERR_INVALID, "null pointer passed to pgx_plan_build_bitmap");
    const long long sb_rows = 1024ll * slice_words, line = 32ll * slice_words;
    const long long n_super = (n_long + sb_rows - 1) / sb_rows;
    std::atomic<long long> next{0};
    std::atomic<int> bad{0};
    auto work = [&]() {
        for (;;) {
            const long long sb = next.fetch_add(1);
            if (sb >= n_super) return;
            uint32_t *base = bits + static_cast<size_t>(sb) * n_genomes * line;
            const long long r1 = std::min<long long>(n_long, (sb + 1) * sb_rows);
            for (long long r = sb * sb_rows; r < r1; ++r) {
                const long long local = r - sb * sb_rows, gene = long_gene[r];
                const uint32_t bit = 1u << (local & 31);
                const long long word = local >> 5;
                for (long long e = indptr[gene]; e < indptr[gene + 1]; ++e) {
                    const int32_t c = indices[e];
                    if (c < 0 || c >= n_genomes) {
                        bad.store(1);
                        continue;
                    }
                    base[static_cast<size_t>(c) * line + word] |= bit;
                }
            }
        }
    };
    int threads = n_threads > 0 ? n_threads : static_cast<int>(std::min(16u, std::max(1u, std::thread::hardware_concurrency())));
    threads = static_cast<int>(std::min<long long>(threads, n_super));
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work);
    work();
    for (auto &th : pool) th.join();
    if (bad.load()) return pgx::fail(PGX_ERR_INVALID, "genome index out of range in pgx_plan_build_bitmap");
    return PGX_OK;
}
