// Host-side numpy-legacy permutation stream of libpgx_b200 (plain C++, no CUDA).
//
// Replaces the pair ``shuffle_indices = np.arange(N); np.random.shuffle(shuffle_indices)`` of
// /root/reference/pangenomix/pangenome_analysis.py:84-85 for the legacy global RandomState:
// MT19937 (Matsumoto & Nishimura) + Fisher-Yates from the top, j in [0, i] by masked
// rejection on successive 32-bit outputs (mask = smallest 2^b - 1 >= i), bit-exactly.
//
// The stream is serial, but only two of its three stages are: (1) generating raw MT19937
// words and (2) deciding which words are accepted (that decides where the next shuffle
// starts).  Both are vectorised here (AVX-512 or AVX2 when the CPU has it).  Stage (3), applying the
// accepted swap targets to the identity permutation, is independent per shuffle and runs
// on worker threads, two shuffles interleaved per thread for instruction-level parallelism.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#define PGX_X86 1
#else
#define PGX_X86 0
#endif

#include "pgx.h"

namespace pgx {
int fail(int code, const char *fmt, ...);
}

namespace {

constexpr uint32_t MT_UPPER = 0x80000000u, MT_LOWER = 0x7fffffffu, MT_MAGIC = 0x9908b0dfu;

// ``pos`` counts the words of the current block already consumed (624 = block exhausted),
// exactly as np.random.get_state() reports it.
struct Mt19937 {
    alignas(64) uint32_t key[624 + 8];
    alignas(64) uint32_t prev_key[624 + 8];        // key of the block the carried words came from (carry_tail)
    alignas(64) uint32_t out_store[32 + 624 + 8];
    uint32_t *const out = out_store + 32;           // 64-byte aligned; out[-32 .. 0) holds a carried block tail
    int pos;                                        // next word: 0 .. 624, or negative while carried words remain
    bool avx2;
    bool avx512 = false;
    bool bmi2 = false;
    bool carried = false;                           // the current block was entered through carry_tail

    static inline uint32_t twist(uint32_t a, uint32_t b, uint32_t far)
    {
        const uint32_t y = (a & MT_UPPER) | (b & MT_LOWER);
        return far ^ (y >> 1) ^ (-(y & 1u) & MT_MAGIC);
    }

    static inline uint32_t temper1(uint32_t y)
    {
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }

    void temper_scalar()
    {
        for (int k = 0; k < 624; ++k) out[k] = temper1(key[k]);
    }

    void refill_scalar()
    {
        int k = 0;
        for (; k < 624 - 397; ++k) key[k] = twist(key[k], key[k + 1], key[k + 397]);
        for (; k < 623; ++k) key[k] = twist(key[k], key[k + 1], key[k + (397 - 624)]);
        key[623] = twist(key[623], key[0], key[396]);
        temper_scalar();
    }

#if PGX_X86
    __attribute__((target("avx2"))) static inline __m256i twist8(__m256i a, __m256i b, __m256i far)
    {
        const __m256i y = _mm256_or_si256(_mm256_and_si256(a, _mm256_set1_epi32(static_cast<int>(MT_UPPER))),
                                          _mm256_and_si256(b, _mm256_set1_epi32(static_cast<int>(MT_LOWER))));
        const __m256i odd = _mm256_sub_epi32(_mm256_setzero_si256(), _mm256_and_si256(y, _mm256_set1_epi32(1)));
        return _mm256_xor_si256(_mm256_xor_si256(far, _mm256_srli_epi32(y, 1)),
                                _mm256_and_si256(odd, _mm256_set1_epi32(static_cast<int>(MT_MAGIC))));
    }

    __attribute__((target("avx2"))) void refill_avx2()
    {
        // key[k] for k < 227 needs OLD key[k + 1], key[k + 397]: reads run ahead of writes.
        int k = 0;
        for (; k + 8 <= 227; k += 8) {
            const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(key + k));
            const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(key + k + 1));
            const __m256i f = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(key + k + 397));
            _mm256_storeu_si256(reinterpret_cast<__m256i *>(key + k), twist8(a, b, f));
        }
        for (; k < 227; ++k) key[k] = twist(key[k], key[k + 1], key[k + 397]);
        // 227 <= k < 623 needs NEW key[k - 227] (written >= 227 steps ago) and OLD key[k + 1]
        for (; k + 8 <= 623; k += 8) {
            const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(key + k));
            const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(key + k + 1));
            const __m256i f = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(key + k - 227));
            _mm256_storeu_si256(reinterpret_cast<__m256i *>(key + k), twist8(a, b, f));
        }
        for (; k < 623; ++k) key[k] = twist(key[k], key[k + 1], key[k - 227]);
        key[623] = twist(key[623], key[0], key[396]);
        for (k = 0; k < 624; k += 8) {
            __m256i y = _mm256_load_si256(reinterpret_cast<const __m256i *>(key + k));
            y = _mm256_xor_si256(y, _mm256_srli_epi32(y, 11));
            y = _mm256_xor_si256(y, _mm256_and_si256(_mm256_slli_epi32(y, 7), _mm256_set1_epi32(static_cast<int>(0x9d2c5680u))));
            y = _mm256_xor_si256(y, _mm256_and_si256(_mm256_slli_epi32(y, 15), _mm256_set1_epi32(static_cast<int>(0xefc60000u))));
            y = _mm256_xor_si256(y, _mm256_srli_epi32(y, 18));
            _mm256_store_si256(reinterpret_cast<__m256i *>(out + k), y);
        }
    }
#endif

#if PGX_X86
    __attribute__((target("avx512f"))) static inline __m512i twist16(__m512i a, __m512i b, __m512i far)
    {
        const __m512i y = _mm512_or_si512(_mm512_and_si512(a, _mm512_set1_epi32(static_cast<int>(MT_UPPER))),
                                          _mm512_and_si512(b, _mm512_set1_epi32(static_cast<int>(MT_LOWER))));
        const __m512i odd = _mm512_sub_epi32(_mm512_setzero_si512(), _mm512_and_si512(y, _mm512_set1_epi32(1)));
        return _mm512_xor_si512(_mm512_xor_si512(far, _mm512_srli_epi32(y, 1)),
                                _mm512_and_si512(odd, _mm512_set1_epi32(static_cast<int>(MT_MAGIC))));
    }

    __attribute__((target("avx512f"))) void refill_avx512()
    {
        int k = 0;
        for (; k + 16 <= 227; k += 16) {
            const __m512i a = _mm512_loadu_si512(key + k), b = _mm512_loadu_si512(key + k + 1);
            const __m512i f = _mm512_loadu_si512(key + k + 397);
            _mm512_storeu_si512(key + k, twist16(a, b, f));
        }
        for (; k < 227; ++k) key[k] = twist(key[k], key[k + 1], key[k + 397]);
        for (; k + 16 <= 623; k += 16) {
            const __m512i a = _mm512_loadu_si512(key + k), b = _mm512_loadu_si512(key + k + 1);
            const __m512i f = _mm512_loadu_si512(key + k - 227);
            _mm512_storeu_si512(key + k, twist16(a, b, f));
        }
        for (; k < 623; ++k) key[k] = twist(key[k], key[k + 1], key[k - 227]);
        key[623] = twist(key[623], key[0], key[396]);
        for (k = 0; k < 624; k += 16) {
            __m512i y = _mm512_load_si512(key + k);
            y = _mm512_xor_si512(y, _mm512_srli_epi32(y, 11));
            y = _mm512_xor_si512(y, _mm512_and_si512(_mm512_slli_epi32(y, 7), _mm512_set1_epi32(static_cast<int>(0x9d2c5680u))));
            y = _mm512_xor_si512(y, _mm512_and_si512(_mm512_slli_epi32(y, 15), _mm512_set1_epi32(static_cast<int>(0xefc60000u))));
            y = _mm512_xor_si512(y, _mm512_srli_epi32(y, 18));
            _mm512_store_si512(out + k, y);
        }
    }
#endif

    // The last 624 - pos (< 32) words of the block move in front of the next block, so that vector loads
    // run across the block boundary; the old key is kept in case the call ends inside the carried words
    // (the numpy state to report is then {prev_key, 624 - remaining}).
    void carry_tail()
    {
        const int rem = 624 - pos;
        uint32_t tail[32];
        memcpy(tail, out + pos, sizeof(uint32_t) * rem);
        memcpy(prev_key, key, sizeof(uint32_t) * 624);
        refill();
        memcpy(out - rem, tail, sizeof(uint32_t) * rem);
        pos = -rem;
        carried = true;
    }

    void refill()
    {
        carried = false;
#if PGX_X86
        if (avx512) {
            refill_avx512();
            pos = 0;
            return;
        }
        if (avx2) {
            refill_avx2();
            pos = 0;
            return;
        }
#endif
        refill_scalar();
        pos = 0;
    }
};

// Compress table for AVX2: lane permutation that packs the lanes selected by an 8-bit mask.
struct CompressLut {
    alignas(32) uint32_t idx[256][8];
    CompressLut()
    {
        for (int m = 0; m < 256; ++m) {
            int w = 0;
            for (int l = 0; l < 8; ++l)
                if (m & (1 << l)) idx[m][w++] = l;
            for (; w < 8; ++w) idx[m][w] = 0;
        }
    }
};
const CompressLut g_lut;

// Stage 2: the accepted swap targets of ONE shuffle of n elements: js[t] = j drawn for
// i = n - 1 - t.  js must have room for n - 1 + 8 entries (vector stores overrun).
void accept_scalar_span(Mt19937 &mt, uint32_t &i, uint32_t lo, uint32_t mask, uint32_t *&w)
{
    while (i >= lo) {
        if (mt.pos >= 624) mt.refill();
        const uint32_t *src = mt.out + mt.pos;
        const int avail = 624 - mt.pos;
        int t = 0;
        for (; t < avail && i >= lo; ++t) {
            const uint32_t v = src[t] & mask;
            const uint32_t ok = v <= i;
            *w = v;
            w += ok;
            i -= ok;
        }
        mt.pos += t;
    }
}

#if PGX_X86
__attribute__((target("avx2,popcnt")))
void accept_avx2_span(Mt19937 &mt, uint32_t &i_io, uint32_t lo, uint32_t mask, uint32_t *&w_io)
{
    uint32_t i = i_io;
    uint32_t *w = w_io;
    int pos = mt.pos;
    const __m256i maskv = _mm256_set1_epi32(static_cast<int>(mask));
    // all 8 draws of a vector see the same mask as long as i - 7 >= lo even if every one is accepted
    while (i >= lo + 8) {
        if (pos + 8 > 624) {
            if (pos >= 624) { mt.refill(); pos = 0; continue; }
            break;                                   // block tail: the scalar loop finishes it
        }
        const __m256i v = _mm256_and_si256(_mm256_loadu_si256(reinterpret_cast<const __m256i *>(mt.out + pos)), maskv);
        // lane l is tested against i - (accepts among lanes < l), which lies in [i - 7, i]
        const __m256i iv = _mm256_set1_epi32(static_cast<int>(i));
        const __m256i sure = _mm256_cmpgt_epi32(_mm256_sub_epi32(iv, _mm256_set1_epi32(6)), v);      // v <= i - 7
        const __m256i over = _mm256_cmpgt_epi32(v, iv);                                              // v >  i
        const int sure_m = _mm256_movemask_ps(_mm256_castsi256_ps(sure));
        const int both_m = _mm256_movemask_ps(_mm256_castsi256_ps(_mm256_or_si256(sure, over)));
        if (both_m != 0xff) break;                   // an ambiguous lane: resolve this vector one by one
        const __m256i packed = _mm256_permutevar8x32_epi32(v, _mm256_load_si256(reinterpret_cast<const __m256i *>(g_lut.idx[sure_m])));
        _mm256_storeu_si256(reinterpret_cast<__m256i *>(w), packed);
        const uint32_t got = static_cast<uint32_t>(_mm_popcnt_u32(static_cast<unsigned>(sure_m)));
        w += got;
        i -= got;
        pos += 8;
    }
    mt.pos = pos;
    i_io = i;
    w_io = w;
}
#endif

#if PGX_X86
__attribute__((target("avx512f,popcnt")))
void accept_avx512_span(Mt19937 &mt, uint32_t &i_io, uint32_t lo, uint32_t mask, uint32_t *&w_io)
{
    uint32_t i = i_io;
    uint32_t *w = w_io;
    int pos = mt.pos;
    const __m512i maskv = _mm512_set1_epi32(static_cast<int>(mask));
    // The 16-draw steps that finish a mask regime after accept_avx512_span32 (fewer than 32 accepts left):
    // v <= i - 15 is a sure accept, v > i a sure reject, the lanes in between are settled one by one.
    while (i >= lo + 16) {
        if (pos + 16 > 624) {
            mt.pos = pos;
            if (pos >= 624) mt.refill(); else mt.carry_tail();
            pos = mt.pos;
            continue;
        }
        const __m512i v0 = _mm512_and_si512(_mm512_loadu_si512(mt.out + pos), maskv);
        const __m512i top_v = _mm512_set1_epi32(static_cast<int>(i));
        const __mmask16 over0 = _mm512_cmpgt_epu32_mask(v0, top_v);
        unsigned acc = _mm512_cmple_epu32_mask(v0, _mm512_set1_epi32(static_cast<int>(i - 15)));
        unsigned doubt = ~(acc | over0) & 0xffffu;
        if (doubt) {
            alignas(64) uint32_t lanes[16];
            _mm512_store_si512(lanes, v0);
            do {
                const int l = __builtin_ctz(doubt);
                doubt &= doubt - 1;
                const uint32_t before = static_cast<uint32_t>(_mm_popcnt_u32(acc & ((1u << l) - 1u)));
                if (lanes[l] <= i - before) acc |= 1u << l;
            } while (doubt);
        }
        _mm512_storeu_si512(w, _mm512_maskz_compress_epi32(static_cast<__mmask16>(acc), v0));
        const uint32_t got = static_cast<uint32_t>(_mm_popcnt_u32(acc));
        w += got;
        i -= got;
        pos += 16;
    }
    mt.pos = pos;
    i_io = i;
    w_io = w;
}
#endif

#if PGX_X86
// 32 draws per step, every step, while at least 32 accepts remain under this mask.  Lane l is tested against
// i - (accepts before it in the step), which lies in [i - l, i]: v <= i - l is a sure accept, v > i a sure
// reject (two compares against per-lane floors), and the few lanes in between are settled in place, in order,
// from the accepts before them -- no fallback to narrower steps, so the loop-carried chain
// (i -> compares -> popcount -> i) is paid once per 32 draws (10.6-11.0 -> 9.2 us per 10,000-genome shuffle on
// the build host).
__attribute__((target("avx512f,popcnt")))
void accept_avx512_span32(Mt19937 &mt, uint32_t &i_io, uint32_t lo, uint32_t mask, uint32_t *&w_io)
{
    uint32_t i = i_io;
    uint32_t *w = w_io;
    int pos = mt.pos;
    const __m512i maskv = _mm512_set1_epi32(static_cast<int>(mask));
    const __m512i iota0 = _mm512_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    const __m512i iota1 = _mm512_add_epi32(iota0, _mm512_set1_epi32(16));
    while (i >= lo + 32) {
        if (pos + 32 > 624) {
            mt.pos = pos;
            if (pos >= 624) mt.refill(); else mt.carry_tail();
            pos = mt.pos;
            continue;
        }
        const __m512i v0 = _mm512_and_si512(_mm512_loadu_si512(mt.out + pos), maskv);
        const __m512i v1 = _mm512_and_si512(_mm512_loadu_si512(mt.out + pos + 16), maskv);
        const __m512i top_v = _mm512_set1_epi32(static_cast<int>(i));
        // lane l is tested against i - (accepts before it) >= i - l
        const __m512i floor0 = _mm512_sub_epi32(top_v, iota0), floor1 = _mm512_sub_epi32(top_v, iota1);
        const unsigned over = _mm512_cmpgt_epu32_mask(v0, top_v) | (static_cast<unsigned>(_mm512_cmpgt_epu32_mask(v1, top_v)) << 16);
        unsigned acc = _mm512_cmple_epu32_mask(v0, floor0) | (static_cast<unsigned>(_mm512_cmple_epu32_mask(v1, floor1)) << 16);
        unsigned doubt = ~(acc | over);
        if (__builtin_expect(doubt != 0, 0)) {
            alignas(64) uint32_t lanes[32];
            _mm512_store_si512(lanes, v0);
            _mm512_store_si512(lanes + 16, v1);
            do {
                const int l = __builtin_ctz(doubt);
                doubt &= doubt - 1;
                const uint32_t before = static_cast<uint32_t>(_mm_popcnt_u32(acc & ((1u << l) - 1u)));
                if (lanes[l] <= i - before) acc |= 1u << l;
            } while (doubt);
        }
        const uint32_t got0 = static_cast<uint32_t>(_mm_popcnt_u32(acc & 0xffffu));
        const uint32_t got = static_cast<uint32_t>(_mm_popcnt_u32(acc));
        _mm512_storeu_si512(w, _mm512_maskz_compress_epi32(static_cast<__mmask16>(acc), v0));
        _mm512_storeu_si512(w + got0, _mm512_maskz_compress_epi32(static_cast<__mmask16>(acc >> 16), v1));
        w += got;
        i -= got;
        pos += 32;
    }
    mt.pos = pos;
    i_io = i;
    w_io = w;
}
#endif

#if PGX_X86
// The end of a mask regime (fewer than 32 accepts left under this mask, down to the last one) and the small
// regimes at the end of a shuffle, 16 draws per step: the lanes are settled as in accept_avx512_span32, then
// the step is cut behind the accept that ends the regime (its position is the r-th set bit of the accept mask:
// PDEP) and only the draws up to there are consumed -- the next regime starts on the very next draw with its
// own mask.  Replaces the draw-by-draw scalar loop for the ~250 draws per shuffle that no full step covers.
__attribute__((target("avx512f,popcnt,bmi2")))
void accept_avx512_tail(Mt19937 &mt, uint32_t &i_io, uint32_t lo, uint32_t mask, uint32_t *&w_io)
{
    uint32_t i = i_io;
    uint32_t *w = w_io;
    int pos = mt.pos;
    const __m512i maskv = _mm512_set1_epi32(static_cast<int>(mask));
    const __m512i iota = _mm512_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    while (i >= lo) {
        if (pos + 16 > 624) {
            mt.pos = pos;
            if (pos >= 624) mt.refill(); else mt.carry_tail();
            pos = mt.pos;
            continue;
        }
        const __m512i v = _mm512_and_si512(_mm512_loadu_si512(mt.out + pos), maskv);
        const __m512i top_v = _mm512_set1_epi32(static_cast<int>(i));
        // lane l is tested against i - (accepts before it) >= max(i - l, lo) > 0 as long as it lies before the cut
        const __m512i floor_v = _mm512_sub_epi32(_mm512_max_epu32(top_v, iota), iota);
        const unsigned over = _mm512_cmpgt_epu32_mask(v, top_v);
        unsigned acc = _mm512_cmple_epu32_mask(v, floor_v);
        unsigned doubt = ~(acc | over) & 0xffffu;
        if (doubt) {
            alignas(64) uint32_t lanes[16];
            _mm512_store_si512(lanes, v);
            do {
                const int l = __builtin_ctz(doubt);
                doubt &= doubt - 1;
                const uint32_t before = static_cast<uint32_t>(_mm_popcnt_u32(acc & ((1u << l) - 1u)));
                // (a lane behind the cut may see before > i: whatever it decides is cut away below)
                if (lanes[l] <= i - before) acc |= 1u << l;
            } while (doubt);
        }
        const uint32_t left = i - lo + 1;                          // accepts left under this mask, >= 1
        int use = 16;
        if (static_cast<uint32_t>(_mm_popcnt_u32(acc)) >= left) {
            const int last = __builtin_ctz(_pdep_u32(1u << (left - 1), acc));
            use = last + 1;
            acc &= (2u << last) - 1u;
        }
        _mm512_storeu_si512(w, _mm512_maskz_compress_epi32(static_cast<__mmask16>(acc), v));
        const uint32_t got = static_cast<uint32_t>(_mm_popcnt_u32(acc));
        w += got;
        i -= got;
        pos += use;
    }
    mt.pos = pos;
    i_io = i;
    w_io = w;
}
#endif

void accept_one(Mt19937 &mt, uint32_t n, uint32_t *js)
{
    if (n < 2) return;
    uint32_t i = n - 1;
    uint32_t *w = js;
    while (i > 0) {
        const int bits = 32 - __builtin_clz(i);
        const uint32_t mask = 0xffffffffu >> (32 - bits);          // smallest 2^b - 1 >= i
        const uint32_t lo = 1u << (bits - 1);                      // the mask holds while i >= lo
#if PGX_X86
        if (mt.avx512 && mt.bmi2 && mask < 0x80000000u) {
            accept_avx512_span32(mt, i, lo, mask, w);              // 32 draws per step while >= 32 accepts remain
            accept_avx512_tail(mt, i, lo, mask, w);                // ... and the end of the regime, cut at its last accept
            continue;
        }
        if (mt.avx2 && mask < 0x80000000u) {
            // alternate: vector spans while far from lo and unambiguous, scalar for what is left
            const uint32_t stretch = mt.avx512 ? 16 : 8;
            while (i >= lo) {
                if (mt.avx512) {
                    accept_avx512_span32(mt, i, lo, mask, w);
                    accept_avx512_span(mt, i, lo, mask, w);       // fewer than 32 accepts left under this mask
                }
                else accept_avx2_span(mt, i, lo, mask, w);
                if (i < lo) break;
                // one scalar draw-by-draw stretch of at most a vector's worth of draws, then try vectors again
                uint32_t done = 0;
                while (i >= lo && done < stretch) {
                    if (mt.pos >= 624) mt.refill();
                    const uint32_t v = mt.out[mt.pos++] & mask;
                    const uint32_t ok = v <= i;
                    *w = v;
                    w += ok;
                    i -= ok;
                    ++done;
                }
            }
            continue;
        }
#endif
        accept_scalar_span(mt, i, lo, mask, w);
    }
}

// Stage 3: Fisher-Yates from the top with the accepted targets; two shuffles at a time.
inline void apply_one(uint32_t n, const uint32_t *js, uint16_t *a)
{
    for (uint32_t k = 0; k < n; ++k) a[k] = static_cast<uint16_t>(k);
    for (uint32_t i = n - 1, t = 0; n > 1 && i > 0; --i, ++t) {
        const uint32_t j = js[t];
        const uint16_t ai = a[i], aj = a[j];
        a[i] = aj;
        a[j] = ai;
    }
}

inline void apply_two(uint32_t n, const uint32_t *js0, uint16_t *a0, const uint32_t *js1, uint16_t *a1)
{
    for (uint32_t k = 0; k < n; ++k) a0[k] = a1[k] = static_cast<uint16_t>(k);
    for (uint32_t i = n - 1, t = 0; n > 1 && i > 0; --i, ++t) {
        const uint32_t j0 = js0[t], j1 = js1[t];
        const uint16_t x0 = a0[i], y0 = a0[j0];
        const uint16_t x1 = a1[i], y1 = a1[j1];
        a0[i] = y0;
        a0[j0] = x0;
        a1[i] = y1;
        a1[j1] = x1;
    }
}

void apply_batch(uint32_t n, size_t stride, const uint32_t *js, uint16_t *perms, int count)
{
    int s = 0;
    for (; s + 2 <= count; s += 2)
        apply_two(n, js + s * stride, perms + static_cast<size_t>(s) * n, js + (s + 1) * stride,
                  perms + static_cast<size_t>(s + 1) * n);
    if (s < count) apply_one(n, js + s * stride, perms + static_cast<size_t>(s) * n);
}

int worker_threads_wanted()
{
    if (const char *env = getenv("PGX_RNG_THREADS")) {
        const int v = atoi(env);
        if (v >= 0) return std::min(v, 64);
    }
    const unsigned hw = std::thread::hardware_concurrency();
    return static_cast<int>(std::min(4u, hw > 1 ? hw - 1 : 0u));
}

// Stage-3 workers, created on first use and kept (see pgx_legacy_shuffles).  Batches are handed over through
// atomics, not a condition variable: waking a sleeping thread costs 0.2-0.5 ms on the virtualised hosts this
// runs on (measured: 17 submits of a 209-shuffle call took 8 ms, three times the acceptance work itself), so
// idle workers spin for a few milliseconds -- long enough to stay awake from one block-sized call to the next
// -- before they go to sleep, and the producer only pays for a wake-up when somebody is in fact asleep.
// The object is leaked on purpose: at process exit its threads must not see it destroyed.
class WorkerPool {
public:
    static constexpr int MAX_THREADS = 16;
    static constexpr int MAX_SLOTS = 2 * MAX_THREADS + 2;

    static WorkerPool &get(int threads)
    {
        static WorkerPool *pool = nullptr;
        if (!pool || pool->pid_ != getpid()) pool = new WorkerPool();     // a forked child has no threads: start over
        pool->grow(threads);
        return *pool;
    }

    // Call context; only the producer calls begin / slot_for_next / submit / finish, one call at a time, and
    // every batch of the previous call is complete when begin runs.
    void begin(uint32_t n, size_t stride, uint32_t *ring, int batch, int n_slots, uint16_t *perms, int active)
    {
        n_ = n;
        stride_ = stride;
        ring_ = ring;
        batch_ = batch;
        perms_ = perms;
        n_slots_ = n_slots;
        base_ = produced_.load(std::memory_order_relaxed);
        for (int s = 0; s < n_slots; ++s) slot_done_[s].store(base_, std::memory_order_relaxed);
        active_.store(active, std::memory_order_release);
    }

    // Buffer of the next batch: its slot is free once the batch that used it n_slots batches ago is complete.
    int slot_for_next()
    {
        const int64_t id = produced_.load(std::memory_order_relaxed);
        const int s = static_cast<int>((id - base_) % n_slots_);
        while (slot_done_[s].load(std::memory_order_acquire) < id - n_slots_ + 1) {
            if (!run_one()) cpu_relax();                                   // the producer helps instead of waiting
        }
        return s;
    }

    void submit(int slot, int64_t first, int count)
    {
        jobs_[slot].first = first;
        jobs_[slot].count = count;
        produced_.fetch_add(1, std::memory_order_seq_cst);                 // publishes the job and the call context
        if (sleepers_.load(std::memory_order_seq_cst) > 0) {
            { std::lock_guard<std::mutex> lock(mu_); }
            cv_.notify_all();
        }
    }

    void finish()
    {
        while (run_one()) {}
        const int64_t all = produced_.load(std::memory_order_relaxed);
        while (completed_.load(std::memory_order_acquire) < all) cpu_relax();
    }

private:
    struct Job { int64_t first = 0; int count = 0; };

    WorkerPool() : pid_(getpid()) {}

    static inline void cpu_relax()
    {
#if PGX_X86
        _mm_pause();
#endif
    }

    void grow(int threads)
    {
        threads = std::min(threads, MAX_THREADS);
        while (n_threads_ < threads) {
            const int idx = n_threads_++;
            std::thread([this, idx] { work(idx); }).detach();
        }
    }

    // Claims and applies one submitted batch; false if none is waiting.
    bool run_one()
    {
        int64_t id = claimed_.load(std::memory_order_relaxed);
        for (;;) {
            if (id >= produced_.load(std::memory_order_acquire)) return false;
            if (claimed_.compare_exchange_weak(id, id + 1, std::memory_order_acq_rel, std::memory_order_relaxed)) break;
        }
        const int s = static_cast<int>((id - base_) % n_slots_);
        const Job job = jobs_[s];
        apply_batch(n_, stride_, ring_ + stride_ * batch_ * s, perms_ + static_cast<size_t>(job.first) * n_, job.count);
        slot_done_[s].store(id + 1, std::memory_order_release);
        completed_.fetch_add(1, std::memory_order_release);
        return true;
    }

    void work(int idx)
    {
        auto last_work = std::chrono::steady_clock::now();
        unsigned rounds = 0;
        for (;;) {
            if (idx < active_.load(std::memory_order_acquire) && run_one()) {
                last_work = std::chrono::steady_clock::now();
                continue;
            }
            cpu_relax();
            if ((++rounds & 255u) || std::chrono::steady_clock::now() - last_work < std::chrono::milliseconds(spin_ms_)) continue;
            // nothing for a few milliseconds: sleep until the producer submits again (the timeout is a safety net)
            {
                std::unique_lock<std::mutex> lock(mu_);
                sleepers_.fetch_add(1, std::memory_order_seq_cst);
                if (claimed_.load(std::memory_order_seq_cst) >= produced_.load(std::memory_order_seq_cst) ||
                    idx >= active_.load(std::memory_order_acquire))
                    cv_.wait_for(lock, std::chrono::milliseconds(200));
                sleepers_.fetch_sub(1, std::memory_order_seq_cst);
            }
            last_work = std::chrono::steady_clock::now();
        }
    }

    // idle spinning before a worker goes to sleep (PGX_RNG_SPIN_MS; 0 = sleep at once and pay for every wake-up)
    const int spin_ms_ = getenv("PGX_RNG_SPIN_MS") ? std::max(0, atoi(getenv("PGX_RNG_SPIN_MS"))) : 3;

    // call context (plain: published by the first submit of the call, read only after a claim)
    uint32_t n_ = 0;
    size_t stride_ = 0;
    uint32_t *ring_ = nullptr;
    uint16_t *perms_ = nullptr;
    int batch_ = 0, n_slots_ = 1, n_threads_ = 0;
    int64_t base_ = 0;
    Job jobs_[MAX_SLOTS];
    std::atomic<int64_t> slot_done_[MAX_SLOTS] = {};
    std::atomic<int64_t> produced_{0}, claimed_{0}, completed_{0};
    std::atomic<int> active_{0}, sleepers_{0};
    std::mutex mu_;
    std::condition_variable cv_;
    const pid_t pid_;
};

}  // namespace

extern "C" int pgx_legacy_shuffles(uint32_t *mt_key, int32_t *mt_pos, int64_t n, int64_t count,
                                   uint16_t *h_perms)
{
    if (!mt_key || !mt_pos || (!h_perms && n * count > 0))
        return pgx::fail(PGX_ERR_INVALID, "null pointer passed to pgx_legacy_shuffles");
    if (n < 0 || n > 65535) return pgx::fail(PGX_ERR_UNSUPPORTED, "n = %lld outside 0..65535", (long long)n);
    if (count < 0 || *mt_pos < 0 || *mt_pos > 624) return pgx::fail(PGX_ERR_INVALID, "bad MT19937 position / count");

    static Mt19937 mt_storage;                 // 5 KB; calls are serialised by the caller's RNG ownership anyway
    static std::mutex call_mu;
    std::lock_guard<std::mutex> call_lock(call_mu);
    Mt19937 &mt = mt_storage;
    memcpy(mt.key, mt_key, sizeof(uint32_t) * 624);
    mt.pos = *mt_pos;
    mt.carried = false;
#if PGX_X86
    mt.avx2 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("popcnt") && !getenv("PGX_RNG_SCALAR");
    mt.avx512 = mt.avx2 && __builtin_cpu_supports("avx512f") && !getenv("PGX_RNG_NO_AVX512");
    mt.bmi2 = mt.avx512 && __builtin_cpu_supports("bmi2") && !getenv("PGX_RNG_NO_TAIL");
#else
    mt.avx2 = false;
    mt.avx512 = false;
    mt.bmi2 = false;
#endif
    mt.temper_scalar();

    const uint32_t un = static_cast<uint32_t>(n);
    const size_t stride = static_cast<size_t>(n) + 16;         // js row, with room for vector overrun
    const int batch = static_cast<int>(std::max<int64_t>(2, std::min<int64_t>(64, (1 << 17) / std::max<int64_t>(n, 1))));
    int workers = (n >= 64 && count >= 4 * batch) ? worker_threads_wanted() : 0;

    if (workers == 0) {
        std::vector<uint32_t> js(stride * 2);
        int64_t t = 0;
        for (; t + 2 <= count; t += 2) {
            accept_one(mt, un, js.data());
            accept_one(mt, un, js.data() + stride);
            apply_two(un, js.data(), h_perms + t * n, js.data() + stride, h_perms + (t + 1) * n);
        }
        if (t < count) {
            accept_one(mt, un, js.data());
            apply_one(un, js.data(), h_perms + t * n);
        }
    } else {
        // producer (this thread): stages 1-2 into a ring of batch buffers; workers: stage 3.
        // The ring's buffers and the worker threads outlive the call (guarded by call_mu): a caller that draws
        // block after block (pgx_estimate_pan_core, 209 shuffles per call) would otherwise pay for zero-filling
        // ~5 MB of fresh pages per call and, far worse, for freshly created threads that the scheduler wakes on
        // the producer's own core until its load balancer has spread them (measured on the build host: 38 us
        // per shuffle in calls of 209 against 14 us in one long call, all of it in the producer's loop).
        workers = std::min(workers, WorkerPool::MAX_THREADS);
        WorkerPool &pool = WorkerPool::get(workers);
        const int n_slots = 2 * workers + 2;
        static std::vector<uint32_t> ring;
        if (ring.size() < stride * batch * n_slots) ring.resize(stride * batch * n_slots);
        pool.begin(un, stride, ring.data(), batch, n_slots, h_perms, workers);
        for (int64_t t = 0; t < count; t += batch) {
            const int s = pool.slot_for_next();
            const int c = static_cast<int>(std::min<int64_t>(batch, count - t));
            uint32_t *js = ring.data() + stride * batch * s;
            for (int b = 0; b < c; ++b) accept_one(mt, un, js + b * stride);
            pool.submit(s, t, c);
        }
        pool.finish();                                   // the producer applies what is still queued, then waits
    }
    if (mt.pos < 0 || (mt.pos == 0 && mt.carried)) {   // the call ended inside (or exactly at the end of) a carried block tail:
                                                       // numpy is still on the old block, at position 624 - remaining
        memcpy(mt_key, mt.prev_key, sizeof(uint32_t) * 624);
        *mt_pos = 624 + mt.pos;
    } else {
        memcpy(mt_key, mt.key, sizeof(uint32_t) * 624);
        *mt_pos = mt.pos;
    }
    return PGX_OK;
}

// Raw 32-bit outputs of the same stream: what numpy's legacy ``random_sample`` (two words per double), and through it
// ``RandomState.choice(p=...)`` -- draw_bbn, /root/reference/pangenomix/pangenome_analysis.py:484-492 -- consume.
namespace {
// Sequential writer into a large destination: whole 64-byte lines leave with streaming stores (no read-for-ownership
// of the destination), assembled in a pending line so that neither the block boundaries of the generator (624 words)
// nor the alignment of the destination ever split a line; ordinary stores for the unaligned head and tail.
struct StreamWriter {
    uint32_t *dst;
    alignas(64) uint32_t line[16];
    int pending = 0;                                    // words waiting in ``line`` (dst is 64-byte aligned while > 0)
    bool wide;
    StreamWriter(uint32_t *d, bool avx512) : dst(d), wide(avx512) {}
#if PGX_X86
    __attribute__((target("avx512f"))) static void lines512(uint32_t *d, const uint32_t *src, int64_t lines)
    {
        for (int64_t i = 0; i < lines; ++i)
            _mm512_stream_si512(reinterpret_cast<__m512i *>(d + 16 * i), _mm512_loadu_si512(src + 16 * i));
    }
    static void lines128(uint32_t *d, const uint32_t *src, int64_t lines)
    {
        for (int64_t i = 0; i < 4 * lines; ++i)
            _mm_stream_si128(reinterpret_cast<__m128i *>(d + 4 * i), _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 4 * i)));
    }
#endif
    void put_lines(const uint32_t *src, int64_t lines)
    {
#if PGX_X86
        if (wide) lines512(dst, src, lines);
        else lines128(dst, src, lines);
#else
        memcpy(dst, src, sizeof(uint32_t) * 16 * lines);
#endif
        dst += 16 * lines;
    }
    void push(const uint32_t *src, int64_t n)
    {
        while (n > 0 && pending == 0 && (reinterpret_cast<uintptr_t>(dst) & 63)) {    // head of the destination
            *dst++ = *src++;
            --n;
        }
        if (pending > 0) {
            const int fill = static_cast<int>(std::min<int64_t>(16 - pending, n));
            memcpy(line + pending, src, sizeof(uint32_t) * fill);
            pending += fill;
            src += fill;
            n -= fill;
            if (pending < 16) return;
            put_lines(line, 1);
            pending = 0;
        }
        const int64_t lines = n / 16;
        if (lines > 0) {
            put_lines(src, lines);
            src += 16 * lines;
            n -= 16 * lines;
        }
        if (n > 0) {
            memcpy(line, src, sizeof(uint32_t) * n);
            pending = static_cast<int>(n);
        }
    }
    void finish()
    {
        if (pending > 0) memcpy(dst, line, sizeof(uint32_t) * pending);
#if PGX_X86
        _mm_sfence();                                   // before the caller hands the buffer to a copy engine
#endif
    }
};
}  // namespace

extern "C" int pgx_legacy_random_raw(uint32_t *mt_key, int32_t *mt_pos, int64_t count, uint32_t *h_out)
{
    if (!mt_key || !mt_pos || (!h_out && count > 0))
        return pgx::fail(PGX_ERR_INVALID, "null pointer passed to pgx_legacy_random_raw");
    if (count < 0 || *mt_pos < 0 || *mt_pos > 624) return pgx::fail(PGX_ERR_INVALID, "bad MT19937 position / count");
    static Mt19937 mt_storage;
    static std::mutex call_mu;
    std::lock_guard<std::mutex> call_lock(call_mu);
    Mt19937 &mt = mt_storage;
    memcpy(mt.key, mt_key, sizeof(uint32_t) * 624);
    mt.pos = *mt_pos;
    mt.carried = false;
#if PGX_X86
    mt.avx2 = __builtin_cpu_supports("avx2") && !getenv("PGX_RNG_SCALAR");
    mt.avx512 = mt.avx2 && __builtin_cpu_supports("avx512f") && !getenv("PGX_RNG_NO_AVX512");
#else
    mt.avx2 = false;
    mt.avx512 = false;
#endif
    mt.bmi2 = false;
    if (mt.pos < 624) mt.temper_scalar();
    StreamWriter writer(h_out, mt.avx512);
    int64_t done = 0;
    while (done < count) {
        if (mt.pos == 624) mt.refill();                 // numpy refills lazily too: a call that ends on a block
                                                        // boundary reports position 624 of the old key
        const int64_t take = std::min<int64_t>(624 - mt.pos, count - done);
        writer.push(mt.out + mt.pos, take);
        mt.pos += static_cast<int>(take);
        done += take;
    }
    writer.finish();
    memcpy(mt_key, mt.key, sizeof(uint32_t) * 624);
    *mt_pos = mt.pos;
    return PGX_OK;
}
