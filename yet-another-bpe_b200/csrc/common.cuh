// common.cuh -- shared device helpers for the yabpe sm_100a kernels.
//
// Everything here is integer/byte work: Unicode class lookup,
// UTF-8 decoding, hashing, and the GPT-2 pre-token start rule in its generic
// (global-memory) form.  The tile kernel in pretok.cuh uses a shared-memory
// specialisation of the same rule.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "unicode_tables.inc"


typedef unsigned long long u64;
typedef long long i64;

// class codes used on the device (5 values so that "previous char is U+0020" is a class test)
#define KC_O 0
#define KC_L 1
#define KC_N 2
#define KC_S 3      // \s other than U+0020
#define KC_SP 4     // U+0020

// per-byte info bits (shared by the tile kernel and the generic accessor)
#define IB_CLS 0x07
#define IB_CONT 0x08   // UTF-8 continuation byte
#define IB_INC 0x10    // inside a live contraction -> never a token start
#define IB_FL 0x20     // the logical text starts here (offset 0, chunk cut, end of a recognised special)
#define IB_FR 0x40     // the logical text ends before this byte (chunk cut, end of buffer, encode-mode special start)
#define IB_IN 0x80     // byte of a recognised special token

__device__ unsigned char g_ucd_stage1[4352];
__device__ unsigned char g_ucd_stage2[YABPE_UCD_NBLOCKS * 64];

__device__ __forceinline__ int ucd_class(uint32_t cp) {
    // 0=O 1=L 2=N 3=S ; cp must be < 0x110000
    unsigned blk = g_ucd_stage1[cp >> 8];
    unsigned byte = g_ucd_stage2[blk * 64 + ((cp & 255) >> 2)];
    return (byte >> ((cp & 3) * 2)) & 3;
}

__device__ __forceinline__ int kclass_of_cp(uint32_t cp) {
    int c = ucd_class(cp);
    return (c == 3 && cp == 0x20) ? KC_SP : c;
}

__device__ __forceinline__ u64 mix64(u64 x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

// position-mixed additive hash for long pre-tokens: parallelisable (sum of per-byte terms)
__device__ __forceinline__ u64 long_hash_term(uint8_t b, i64 i) {
    return mix64(((u64)i << 8 | b) + 0x9e3779b97f4a7c15ULL);
}

__device__ __forceinline__ int utf8_len_from_lead(uint8_t b) {
    return b < 0x80 ? 1 : (b < 0xE0 ? 2 : (b < 0xF0 ? 3 : 4));
}

// Strict UTF-8: is the sequence starting at lead byte p[0] well formed?  avail = readable bytes.
__device__ __forceinline__ bool utf8_seq_ok(const uint8_t* p, i64 avail, int* len_out) {
    uint8_t b = p[0];
    if (b < 0x80) { *len_out = 1; return true; }
    int need; uint8_t lo = 0x80, hi = 0xBF;
    if (b >= 0xC2 && b <= 0xDF) need = 1;
    else if (b == 0xE0) { need = 2; lo = 0xA0; }
    else if ((b >= 0xE1 && b <= 0xEC) || b == 0xEE || b == 0xEF) need = 2;
    else if (b == 0xED) { need = 2; hi = 0x9F; }
    else if (b == 0xF0) { need = 3; lo = 0x90; }
    else if (b >= 0xF1 && b <= 0xF3) need = 3;
    else if (b == 0xF4) { need = 3; hi = 0x8F; }
    else { *len_out = 1; return false; }
    *len_out = need + 1;
    if (avail < need + 1) return false;
    if (p[1] < lo || p[1] > hi) return false;
    for (int k = 2; k <= need; k++) if ((p[k] & 0xC0) != 0x80) return false;
    return true;
}

__device__ __forceinline__ uint32_t utf8_decode(const uint8_t* p, int len) {
    if (len == 1) return p[0];
    if (len == 2) return ((uint32_t)(p[0] & 0x1F) << 6) | (p[1] & 0x3F);
    if (len == 3) return ((uint32_t)(p[0] & 0x0F) << 12) | ((uint32_t)(p[1] & 0x3F) << 6) | (p[2] & 0x3F);
    return ((uint32_t)(p[0] & 0x07) << 18) | ((uint32_t)(p[1] & 0x3F) << 12) | ((uint32_t)(p[2] & 0x3F) << 6) | (p[3] & 0x3F);
}

// ---------------------------------------------------------------------------------
// Special tokens live in constant memory (a handful of short strings).
// ---------------------------------------------------------------------------------
#define YABPE_MAX_SPECIALS 64
#define YABPE_MAX_SPECIAL_BYTES 2048

struct SpecialSet {
    int n;
    int max_len;
    int offs[YABPE_MAX_SPECIALS + 1];
    unsigned char blob[YABPE_MAX_SPECIAL_BYTES];
    unsigned char first_byte_mask[32];   // 256-bit set of first bytes
    int n_first;                         // distinct first bytes (<= 8 listed; more -> generic scan)
    unsigned char first[8];
};
__constant__ SpecialSet c_sp;

// first special (priority order) matching at text[i..]; limit = exclusive end of the logical text
__device__ __forceinline__ int special_match(const uint8_t* text, i64 i, i64 limit) {
    uint8_t b = text[i];
    if (!((c_sp.first_byte_mask[b >> 3] >> (b & 7)) & 1)) return -1;
    for (int s = 0; s < c_sp.n; s++) {
        int len = c_sp.offs[s + 1] - c_sp.offs[s];
        if (i + len > limit) continue;
        const unsigned char* q = c_sp.blob + c_sp.offs[s];
        bool ok = true;
        for (int k = 0; k < len; k++) if (text[i + k] != q[k]) { ok = false; break; }
        if (ok) return s;
    }
    return -1;
}

// first hard boundary > i (chunk cut, file / document end, or n): a special never straddles one, so EVERY
// match or re-match of a special at i uses this as its limit -- recognition and the later length look-ups must agree
// even when one special is a prefix of another ('<|eot|>' + cut + 'x' is not '<|eot|>x')
__device__ __forceinline__ i64 hard_end_after(const i64* cuts, int n_cuts, i64 n, i64 i) {
    int lo = 0, hi = n_cuts;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (cuts[mid] <= i) lo = mid + 1; else hi = mid; }
    return lo < n_cuts ? cuts[lo] : n;
}

// ---------------------------------------------------------------------------------
// Generic accessor over global memory.  Fences:
//   cuts[]     sorted hard boundaries (FL + FR at each cut); 0 and n are implicit
//   fence_fl   one extra FL position (end of the previously recognised special), or -1
//   rec bitmap recognised specials (used after resolution; may be null)
// mode 0 = trainer: a recognised special [q, q+m) gives IN on its bytes, FL at q and q+m
// mode 1 = encode : additionally FR at q
// ---------------------------------------------------------------------------------
struct GlobalText {
    const uint8_t* text;
    i64 n;
    const i64* cuts; int n_cuts;
    const uint32_t* rec;     // bitmap of recognised special starts, or null
    i64 fence_fl;            // extra FL (walk state), -1 if none
    int mode;

    __device__ bool is_cut(i64 p) const {
        if (p <= 0) return p == 0;
        if (p >= n) return true;
        int lo = 0, hi = n_cuts;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (cuts[mid] < p) lo = mid + 1; else hi = mid; }
        return lo < n_cuts && cuts[lo] == p;
    }
    __device__ bool rec_bit(i64 q) const { return rec && q >= 0 && q < n && ((rec[q >> 5] >> (q & 31)) & 1); }
    // start q of the recognised special covering byte p (q <= p < q+m), or -1
    __device__ i64 covering_special(i64 p, int* len_out) const {
        if (!rec) return -1;
        i64 lo = p - c_sp.max_len + 1; if (lo < 0) lo = 0;
        for (i64 q = p; q >= lo; q--) {
            if (rec_bit(q)) {
                int s = special_match(text, q, hard_end_after(cuts, n_cuts, n, q));
                int m = s >= 0 ? c_sp.offs[s + 1] - c_sp.offs[s] : 0;
                if (q + m > p) { *len_out = m; return q; }
                return -1;   // recognised specials never overlap: the nearest one decides
            }
        }
        return -1;
    }
    __device__ bool fl(i64 p) const {
        if (p == fence_fl) return true;
        if (is_cut(p)) return true;
        if (rec) {
            if (rec_bit(p)) return true;
            int m; i64 q = p > 0 ? covering_special(p - 1, &m) : -1;
            if (q >= 0 && q + m == p) return true;
        }
        return false;
    }
    __device__ bool fr(i64 p) const {
        if (p >= n) return true;
        if (p > 0 && is_cut(p)) return true;
        if (mode == 1 && rec_bit(p)) return true;
        return false;
    }
    __device__ bool in_special(i64 p) const {   // any byte of a recognised special
        int m; return covering_special(p, &m) >= 0;
    }
    // class code of the code point that byte p belongs to (text assumed valid UTF-8 around p)
    __device__ int kclass(i64 p) const {
        i64 q = p;
        while (q > 0 && p - q < 3 && (text[q] & 0xC0) == 0x80) q--;
        uint8_t b = text[q];
        if (b < 0x80) return kclass_of_cp(b);
        int len = utf8_len_from_lead(b);
        if (q + len > n) return KC_O;
        return kclass_of_cp(utf8_decode(text + q, len));
    }
};

// live contraction starting at apostrophe position a?  returns its length (2 or 3) or 0.
template <class A>
__device__ int live_contraction(const A& t, i64 a) {
    if (t.text[a] != '\'') return 0;
    if (t.in_special(a)) return 0;
    if (t.fr(a + 1)) return 0;
    uint8_t c1 = t.text[a + 1];
    int clen = 0;
    if (c1 == 's' || c1 == 'd' || c1 == 'm' || c1 == 't') clen = 2;
    else if (!t.fr(a + 2)) {
        uint8_t c2 = t.text[a + 2];
        if ((c1 == 'l' && c2 == 'l') || (c1 == 'v' && c2 == 'e') || (c1 == 'r' && c2 == 'e')) clen = 3;
    }
    if (!clen) return 0;
    if (t.fl(a)) return clen;
    int p = t.kclass(a - 1);
    if (p == KC_L || p == KC_N || p == KC_S) return clen;
    return 0;
}

// Is lead-byte position i a pre-token start?  (SURVEY.md Appendix A.1/A.2 restated over bytes.)
template <class A>
__device__ bool is_token_start(const A& t, i64 i) {
    if ((t.text[i] & 0xC0) == 0x80) return false;
    if (t.fl(i)) return true;
    if (t.in_special(i)) return false;
    // contraction context: a live apostrophe at i-1 .. i-3
    for (int d = 1; d <= 3; d++) {
        i64 a = i - d;
        if (a < 0) break;
        uint8_t b = t.text[a];
        if (b == '\'') {
            int cl = live_contraction(t, a);
            if (cl > d) return false;
            if (cl == d) return true;
            break;
        }
        if (b < 'a' || b > 'z') break;
        if (t.fl(a)) break;
    }
    int c = t.kclass(i);
    int p = t.kclass(i - 1);
    if (c < KC_S) {
        if (p == KC_SP) return false;
        if (p >= KC_S) return true;
        return p != c;
    }
    if (p < KC_S) return true;
    i64 nx = i + utf8_len_from_lead(t.text[i]);
    if (t.fr(nx)) return false;
    return t.kclass(nx) < KC_S;
}

__device__ __forceinline__ int warp_reduce_sum(int v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
