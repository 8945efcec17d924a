// Raw DEFLATE (RFC 1951) decoder of libpgx_b200 (plain C++, no CUDA) for the LSDF ingest.
//
// ``read_lsdf`` (/root/reference/pangenomix/sparse_utils.py:18-42 -> scipy.sparse.load_npz) spends its time
// inflating the ``row`` / ``col`` / ``data`` members of the ``.npz`` that ``to_npz`` wrote with
// ``scipy.sparse.save_npz(compressed=True)`` (sparse_utils.py:314): 0.72 GB at C4, one zlib stream per
// member.  This decoder writes a member straight into the numpy array that will hold it: whole-buffer input
// and output, a 64-bit bit buffer refilled with one unaligned load, 11-bit first-level tables with second-level
// tables for longer codes, up to three literals per refill and 8-byte-wide match copies.  The caller checks
// the CRC-32 of the zip entry afterwards and falls back to zlib on any error, so a wrong byte cannot pass.
#include <stdint.h>
#include <string.h>

#include "pgx.h"

namespace pgx {
int fail(int code, const char *fmt, ...);
}

namespace {

constexpr int LITLEN_BITS = 11, DIST_BITS = 8, MAX_CODE_LEN = 15;
constexpr int LITLEN_SYMS = 288, DIST_SYMS = 32, PRECODE_SYMS = 19;
// first-level entries + the most second-level entries a complete code can need (zlib's ENOUGH bounds)
constexpr int LITLEN_TABLE = 2048 + 512, DIST_TABLE = 256 + 512;

// Table entry: bits 0-3 code bits to drop (of this level), bits 4-7 extra bits, bits 8-10 kind, bits 16-31 value.
enum : uint32_t { KIND_LITERAL = 0, KIND_BASE = 1, KIND_END = 2, KIND_SUBTABLE = 3, KIND_INVALID = 4 };
constexpr uint32_t entry(uint32_t kind, uint32_t value, uint32_t extra, uint32_t len)
{
    return (value << 16) | (kind << 8) | (extra << 4) | len;
}
constexpr uint32_t INVALID_ENTRY = entry(KIND_INVALID, 0, 0, 1);

const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073,
                                4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t PRECODE_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

inline uint32_t litlen_entry(int sym, int len)
{
    if (sym < 256) return entry(KIND_LITERAL, static_cast<uint32_t>(sym), 0, len);
    if (sym == 256) return entry(KIND_END, 0, 0, len);
    if (sym <= 285) return entry(KIND_BASE, LEN_BASE[sym - 257], LEN_EXTRA[sym - 257], len);
    return entry(KIND_INVALID, 0, 0, len);                         // 286, 287: in the fixed code, never valid data
}

inline uint32_t dist_entry(int sym, int len)
{
    if (sym < 30) return entry(KIND_BASE, DIST_BASE[sym], DIST_EXTRA[sym], len);
    return entry(KIND_INVALID, 0, 0, len);
}

inline uint32_t reverse_bits(uint32_t code, int len)
{
    uint32_t r = 0;
    for (int b = 0; b < len; ++b) r |= ((code >> b) & 1u) << (len - 1 - b);
    return r;
}

// Canonical Huffman code -> two-level decode table indexed by the next (LSB-first) input bits.
// ``kind``: 0 litlen, 1 distance, 2 precode (value = symbol).  Returns false for an over-subscribed code, or
// an incomplete one (a single distance code of length 1 is allowed, as in zlib); unused slots stay INVALID.
bool build_table(const uint8_t *lens, int n_syms, int table_bits, int which, uint32_t *table, int table_cap)
{
    int count[MAX_CODE_LEN + 1] = {0};
    for (int s = 0; s < n_syms; ++s) ++count[lens[s]];
    count[0] = 0;
    int used = 0, max_len = 0;
    long long kraft = 0;                                           // in units of 2^-15
    for (int l = 1; l <= MAX_CODE_LEN; ++l) {
        used += count[l];
        kraft += static_cast<long long>(count[l]) << (MAX_CODE_LEN - l);
        if (count[l]) max_len = l;
    }
    for (int k = 0; k < (1 << table_bits); ++k) table[k] = INVALID_ENTRY;
    if (used == 0) return which == 1;                              // no distance codes: legal if only literals follow
    if (kraft > (1ll << MAX_CODE_LEN)) return false;
    if (kraft < (1ll << MAX_CODE_LEN) && !(which == 1 && used == 1 && count[1] == 1)) return false;

    uint32_t next_code[MAX_CODE_LEN + 2];
    uint32_t code = 0;
    for (int l = 1; l <= MAX_CODE_LEN; ++l) {
        code = (code + count[l - 1]) << 1;
        next_code[l] = code;
    }
    // second-level tables: for every first-level prefix that long codes share, as many bits as its longest code needs
    uint8_t sub_bits[1 << LITLEN_BITS];
    if (max_len > table_bits) {
        memset(sub_bits, 0, static_cast<size_t>(1) << table_bits);
        uint32_t nc[MAX_CODE_LEN + 2];
        memcpy(nc, next_code, sizeof(nc));
        for (int s = 0; s < n_syms; ++s) {
            const int l = lens[s];
            if (!l) continue;
            const uint32_t rev = reverse_bits(nc[l]++, l);
            if (l > table_bits) {
                uint8_t &b = sub_bits[rev & ((1u << table_bits) - 1u)];
                if (l - table_bits > b) b = static_cast<uint8_t>(l - table_bits);
            }
        }
        int next_free = 1 << table_bits;
        for (int k = 0; k < (1 << table_bits); ++k) {
            if (!sub_bits[k]) continue;
            if (next_free + (1 << sub_bits[k]) > table_cap) return false;
            table[k] = entry(KIND_SUBTABLE, static_cast<uint32_t>(next_free), 0, sub_bits[k]);
            for (int j = 0; j < (1 << sub_bits[k]); ++j) table[next_free + j] = INVALID_ENTRY;
            next_free += 1 << sub_bits[k];
        }
    }
    for (int s = 0; s < n_syms; ++s) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t rev = reverse_bits(next_code[l]++, l);
        if (l <= table_bits) {
            const uint32_t e = which == 0 ? litlen_entry(s, l) : which == 1 ? dist_entry(s, l)
                                                                           : entry(KIND_LITERAL, static_cast<uint32_t>(s), 0, l);
            for (uint32_t k = rev; k < (1u << table_bits); k += 1u << l) table[k] = e;
        } else {
            const uint32_t head = table[rev & ((1u << table_bits) - 1u)];
            const int bits = head & 15;
            const uint32_t base = head >> 16;
            const int rest = l - table_bits;
            const uint32_t e = which == 0 ? litlen_entry(s, rest) : which == 1 ? dist_entry(s, rest)
                                                                              : entry(KIND_LITERAL, static_cast<uint32_t>(s), 0, rest);
            for (uint32_t k = rev >> table_bits; k < (1u << bits); k += 1u << rest) table[base + k] = e;
        }
    }
    return true;
}

struct Inflater {
    const uint8_t *in, *in_end;
    uint8_t *out, *out_begin, *out_end;
    uint64_t bitbuf = 0;
    int bitcnt = 0;                  // valid bits in bitbuf
    long long padded = 0;            // zero bits supplied past the end of the input
    uint32_t litlen[LITLEN_TABLE];
    uint32_t dist[DIST_TABLE];

    inline void refill()
    {
        if (in_end - in >= 8) {
            uint64_t w;
            memcpy(&w, in, 8);
            bitbuf |= w << bitcnt;
            in += (63 - bitcnt) >> 3;
            bitcnt |= 56;
        } else {
            while (bitcnt <= 56) {
                if (in < in_end) bitbuf |= static_cast<uint64_t>(*in++) << bitcnt;
                else padded += 8;
                bitcnt += 8;
            }
        }
    }
    inline uint32_t peek(int n) const { return static_cast<uint32_t>(bitbuf) & ((1u << n) - 1u); }
    inline void drop(int n) { bitbuf >>= n; bitcnt -= n; }
    inline uint32_t take(int n) { const uint32_t v = peek(n); drop(n); return v; }
    // bits consumed beyond the real input: the stream was truncated
    inline bool overran() const { return padded > bitcnt; }

    bool stored_block()
    {
        drop(bitcnt & 7);                                          // to the byte boundary
        refill();
        const uint32_t len = take(16), nlen = take(16);
        if ((len ^ 0xffffu) != nlen || overran()) return false;
        // hand the whole bytes of the bit buffer back to the input
        const int whole = (bitcnt - static_cast<int>(padded > 0 ? (padded < bitcnt ? padded : bitcnt) : 0)) >> 3;
        in -= whole;
        bitbuf = 0;
        bitcnt = 0;
        padded = 0;
        if (static_cast<size_t>(in_end - in) < len || static_cast<size_t>(out_end - out) < len) return false;
        if (len) memcpy(out, in, len);                             // (an empty output may have no buffer at all)
        in += len;
        out += len;
        return true;
    }

    bool read_dynamic_tables()
    {
        refill();
        const int hlit = static_cast<int>(take(5)) + 257, hdist = static_cast<int>(take(5)) + 1, hclen = static_cast<int>(take(4)) + 4;
        if (hlit > 286 || hdist > 30) return false;
        uint8_t pre_lens[PRECODE_SYMS] = {0};
        for (int k = 0; k < hclen; ++k) {
            if (bitcnt < 3) refill();
            pre_lens[PRECODE_ORDER[k]] = static_cast<uint8_t>(take(3));
        }
        uint32_t pre[1 << 7];
        if (!build_table(pre_lens, PRECODE_SYMS, 7, 2, pre, 1 << 7)) return false;
        uint8_t lens[LITLEN_SYMS + DIST_SYMS + 138] = {0};
        int k = 0;
        while (k < hlit + hdist) {
            refill();
            const uint32_t e = pre[peek(7)];
            if (((e >> 8) & 7) != KIND_LITERAL) return false;
            drop(e & 15);
            const int sym = static_cast<int>(e >> 16);
            if (sym < 16) {
                lens[k++] = static_cast<uint8_t>(sym);
            } else {
                int rep;
                uint8_t val = 0;
                if (sym == 16) {
                    if (k == 0) return false;
                    val = lens[k - 1];
                    rep = 3 + static_cast<int>(take(2));
                } else if (sym == 17) {
                    rep = 3 + static_cast<int>(take(3));
                } else {
                    rep = 11 + static_cast<int>(take(7));
                }
                if (k + rep > hlit + hdist) return false;
                memset(lens + k, val, rep);
                k += rep;
            }
            if (overran()) return false;
        }
        if (lens[256] == 0) return false;                          // no end-of-block code
        uint8_t ll[LITLEN_SYMS] = {0}, dl[DIST_SYMS] = {0};
        memcpy(ll, lens, hlit);
        memcpy(dl, lens + hlit, hdist);
        return build_table(ll, LITLEN_SYMS, LITLEN_BITS, 0, litlen, LITLEN_TABLE) &&
               build_table(dl, DIST_SYMS, DIST_BITS, 1, dist, DIST_TABLE);
    }

    bool fixed_tables()
    {
        uint8_t ll[LITLEN_SYMS], dl[DIST_SYMS];
        for (int s = 0; s < 144; ++s) ll[s] = 8;
        for (int s = 144; s < 256; ++s) ll[s] = 9;
        for (int s = 256; s < 280; ++s) ll[s] = 7;
        for (int s = 280; s < 288; ++s) ll[s] = 8;
        for (int s = 0; s < 32; ++s) dl[s] = 5;
        return build_table(ll, LITLEN_SYMS, LITLEN_BITS, 0, litlen, LITLEN_TABLE) &&
               build_table(dl, DIST_SYMS, DIST_BITS, 1, dist, DIST_TABLE);
    }

    // One compressed block.  The fast loop runs while 8 input bytes and FAST_MARGIN output bytes are left, so
    // its refills and wide copies need no bounds checks; the careful loop finishes the block.
    static constexpr long long FAST_MARGIN = 258 + 16 + 3;

    bool compressed_block()
    {
        // The hot state lives in locals: byte stores through ``out`` may alias any member, so members would be
        // reloaded after every store.
        const uint8_t *in_ = in;
        uint8_t *out_ = out;
        uint64_t buf = bitbuf;
        int cnt = bitcnt;
        const uint32_t *const lt = litlen, *const dt = dist;
        const uint8_t *const in_stop = in_end;
        uint8_t *const out_stop = out_end;
        bool ok = false, done = false;
#define PGX_PEEK(n) (static_cast<uint32_t>(buf) & ((1u << (n)) - 1u))
#define PGX_DROP(n) do { buf >>= (n); cnt -= (n); } while (0)
        while (!done) {
            // ---------------- fast loop ----------------
            while (in_stop - in_ >= 8 && out_stop - out_ >= FAST_MARGIN) {
                {                                                  // refill: >= 56 bits = three literals, or length + distance
                    uint64_t w;
                    memcpy(&w, in_, 8);
                    buf |= w << cnt;
                    in_ += (63 - cnt) >> 3;
                    cnt |= 56;
                }
                uint32_t e = lt[PGX_PEEK(LITLEN_BITS)];
                if (__builtin_expect((e & 0x700u) == (KIND_SUBTABLE << 8), 0)) {
                    PGX_DROP(LITLEN_BITS);
                    e = lt[(e >> 16) + PGX_PEEK(e & 15)];
                }
                PGX_DROP(e & 15);
                if ((e & 0x700u) == (KIND_LITERAL << 8)) {
                    *out_++ = static_cast<uint8_t>(e >> 16);
                    // up to two more literals from the same refill (a code is at most 15 bits)
                    e = lt[PGX_PEEK(LITLEN_BITS)];
                    if ((e & 0x700u) != (KIND_LITERAL << 8)) continue;
                    PGX_DROP(e & 15);
                    *out_++ = static_cast<uint8_t>(e >> 16);
                    e = lt[PGX_PEEK(LITLEN_BITS)];
                    if ((e & 0x700u) != (KIND_LITERAL << 8)) continue;
                    PGX_DROP(e & 15);
                    *out_++ = static_cast<uint8_t>(e >> 16);
                    continue;
                }
                if ((e & 0x700u) != (KIND_BASE << 8)) {
                    ok = (e & 0x700u) == (KIND_END << 8);
                    done = true;
                    break;
                }
                const int len_extra = (e >> 4) & 15;
                const uint32_t len = (e >> 16) + PGX_PEEK(len_extra);
                PGX_DROP(len_extra);
                uint32_t d = dt[PGX_PEEK(DIST_BITS)];
                if (__builtin_expect((d & 0x700u) == (KIND_SUBTABLE << 8), 0)) {
                    PGX_DROP(DIST_BITS);
                    d = dt[(d >> 16) + PGX_PEEK(d & 15)];
                }
                if ((d & 0x700u) != (KIND_BASE << 8)) { done = true; break; }
                PGX_DROP(d & 15);
                const int dist_extra = (d >> 4) & 15;
                const uint32_t distance = (d >> 16) + PGX_PEEK(dist_extra);
                PGX_DROP(dist_extra);
                if (distance > static_cast<size_t>(out_ - out_begin)) { done = true; break; }
                const uint8_t *src = out_ - distance;
                uint8_t *dst = out_;
                out_ += len;
                uint8_t *const end = out_;
                if (distance >= 8) {
                    // 16 bytes per turn; may write up to 15 bytes past the match (inside FAST_MARGIN)
                    do {
                        uint64_t w;
                        memcpy(&w, src, 8);
                        memcpy(dst, &w, 8);
                        memcpy(&w, src + 8, 8);
                        memcpy(dst + 8, &w, 8);
                        src += 16;
                        dst += 16;
                    } while (dst < end);
                } else {
                    // distance 1..7: the match repeats a pattern of ``distance`` bytes.  Put as many whole
                    // periods as fit into one 8-byte word and store it every ``stride`` bytes.
                    uint64_t w;
                    uint32_t stride;
                    if (distance == 4) {
                        uint32_t p4;
                        memcpy(&p4, src, 4);
                        w = static_cast<uint64_t>(p4) | (static_cast<uint64_t>(p4) << 32);
                        stride = 8;
                    } else if (distance == 1) {
                        w = 0x0101010101010101ull * src[0];
                        stride = 8;
                    } else if (distance == 2) {
                        uint16_t p2;
                        memcpy(&p2, src, 2);
                        w = 0x0001000100010001ull * p2;
                        stride = 8;
                    } else {
                        uint8_t p[8];
                        for (uint32_t k = 0; k < 8; ++k) p[k] = src[k % distance];
                        memcpy(&w, p, 8);
                        stride = distance == 3 ? 6 : distance;
                    }
                    do {
                        memcpy(dst, &w, 8);
                        dst += stride;
                    } while (dst < end);
                }
            }
            if (done) break;
            // ---------------- careful step: one symbol with every bound checked (ends of the buffers) ----------------
            in = in_; out = out_; bitbuf = buf; bitcnt = cnt;
            const int step = careful_symbol();
            in_ = in; out_ = out; buf = bitbuf; cnt = bitcnt;
            if (step != 0) { ok = step > 0; done = true; }
        }
#undef PGX_PEEK
#undef PGX_DROP
        in = in_; out = out_; bitbuf = buf; bitcnt = cnt;
        return ok;
    }

    // 0: a symbol was decoded, go on; 1: end of block; -1: error.
    int careful_symbol()
    {
        refill();
        uint32_t e = litlen[peek(LITLEN_BITS)];
        if ((e & 0x700u) == (KIND_SUBTABLE << 8)) {
            drop(LITLEN_BITS);
            e = litlen[(e >> 16) + peek(e & 15)];
        }
        drop(e & 15);
        const uint32_t kind = (e >> 8) & 7;
        if (kind == KIND_LITERAL) {
            if (out >= out_end || overran()) return -1;
            *out++ = static_cast<uint8_t>(e >> 16);
            return 0;
        }
        if (kind == KIND_END) return overran() ? -1 : 1;
        if (kind != KIND_BASE) return -1;
        const uint32_t len = (e >> 16) + take((e >> 4) & 15);
        uint32_t d = dist[peek(DIST_BITS)];
        if ((d & 0x700u) == (KIND_SUBTABLE << 8)) {
            drop(DIST_BITS);
            d = dist[(d >> 16) + peek(d & 15)];
        }
        if (((d >> 8) & 7) != KIND_BASE) return -1;
        drop(d & 15);
        const uint32_t distance = (d >> 16) + take((d >> 4) & 15);
        if (overran() || distance > static_cast<size_t>(out - out_begin) || len > static_cast<size_t>(out_end - out)) return -1;
        const uint8_t *src = out - distance;
        for (uint32_t k = 0; k < len; ++k) out[k] = src[k];
        out += len;
        return 0;
    }

    bool run()
    {
        for (;;) {
            refill();
            const uint32_t last = take(1), type = take(2);
            if (overran()) return false;
            bool ok;
            if (type == 0) ok = stored_block();
            else if (type == 1) ok = fixed_tables() && compressed_block();
            else if (type == 2) ok = read_dynamic_tables() && compressed_block();
            else ok = false;
            if (!ok) return false;
            if (last) return out == out_end;
        }
    }
};

}  // namespace

extern "C" int pgx_inflate_raw(const uint8_t *src, int64_t src_len, uint8_t *dst, int64_t dst_len)
{
    if (src_len < 0 || dst_len < 0 || (!src && src_len) || (!dst && dst_len))
        return pgx::fail(PGX_ERR_INVALID, "bad buffer passed to pgx_inflate_raw");
    Inflater state;                 // 13 KB of tables on the stack (a thread_local here costs 25 % in a shared object)
    state.in = src;
    state.in_end = src + src_len;
    state.out = state.out_begin = dst;
    state.out_end = dst + dst_len;
    state.bitbuf = 0;
    state.bitcnt = 0;
    state.padded = 0;
    if (!state.run()) return pgx::fail(PGX_ERR_INVALID, "pgx_inflate_raw: not a DEFLATE stream of exactly %lld bytes", (long long)dst_len);
    return PGX_OK;
}
