// pretok.cuh -- GPU pre-tokeniser (K1) + pre-token counting (K2).
//
// Replaces /root/reference/src/yet_another_bpe/trainer.py:146-170 (process_chunk: strict
// UTF-8 + regex.findall) and trainer.py:221-225 (word_freq), and the pre-tokenisation
// half of tokenizer.py:152-193.  Exact semantics: SURVEY.md section 8(a) rows P1-P5, T2.
//
// Data flow per 8 KiB tile (one CTA, 256 threads, tiles handed out round-robin):
//   TMA bulk copy (cp.async.bulk, double buffered)  HBM -> smem text window
//   pass A  per-byte info (class, continuation, fences), strict UTF-8 validation
//   pass B  live contractions ('s 'd 'm 't 'll 've 're)
//   pass C  token-start bit per byte (warp ballot -> 32-bit masks)
//   pass D  compaction of starts -> one thread per token -> key packing -> hash-table insert
#pragma once

#include "common.cuh"

#ifndef PT_TILE
#define PT_TILE 8192
#endif
#define PT_HL 32
#define PT_HR 256
#define PT_SLACK 48
#define PT_WIN (PT_HL + PT_TILE + PT_HR + PT_SLACK)       // 8528 bytes, multiple of 16
#define PT_THREADS 256
#define PT_NBITS (PT_TILE + PT_HR + 1)                    // start bits for window offsets [HL, HL+TILE+HR]
#define PT_NMASK ((PT_NBITS + 31) / 32)                   // 265 words
#define PT_SHORT_MAX 14
#define PT_MAXMISS 4
#define PT_NSEG (PT_NMASK)                                  // 32-byte segments per tile incl. the right halo

// stats slots (device int64[16])
#define ST_NTOK 0
#define ST_UNIQ_SHORT 1
#define ST_UNIQ_LONG 2
#define ST_UNIQ_BYTES 3
#define ST_ERR_POS 4
#define ST_TABLE_FULL 5
#define ST_OVF_N 6
#define ST_CHAIN_FAIL 7
#define ST_NSPECIAL 8
#define ST_SLOW_N 9        // work items the warp kernel left to the generic tile kernel
#define ST_CACHE_HIT 10    // warp kernel: pre-tokens counted in the shared-memory cache (diagnostic)

struct LongEntry { u64 h; i64 pos; i64 len; i64 count; };
// Short table.  Two layouts behind one accessor pair:
//   interleaved (32-byte slots {k0, k1, count, -}): a probe and the count update that follows it touch ONE DRAM sector --
//               for tables far larger than the L2 (millions of unique pre-tokens)
//   split       (16-byte keys, 8-byte counts in separate arrays): the key lines stay read-only, so probes of hot words
//               do not queue behind the count atomics of the same sector -- for small, heavily contended tables
struct ShortTab {
    char* kb; char* cb; int ks, cs; i64 cap;
    __device__ __forceinline__ u64* key(u64 slot) const { return (u64*)(kb + slot * (u64)ks); }
    __device__ __forceinline__ i64* cnt(u64 slot) const { return (i64*)(cb + slot * (u64)cs); }
};

struct PretokParams {
    const uint8_t* text; i64 n;
    const i64* cuts; int n_cuts;
    int mode; int n_sp;
    i64 own_lo, own_hi;
    uint32_t* cand; uint32_t* rec;
    ShortTab st; i64 scap;
    LongEntry* lent; i64 lcap;
    i64* ovf_pos; i64 ovf_cap;
    i64* stats;
    i64 tile_base; i64 n_tiles;
    const uint4* hot_keys;          // PW_NC keys to pre-load into the warp kernel's cache (from the sizing sample), or null
    ShortTab hot;                   // L2-resident direct-mapped table in front of `st` (interleaved slots; cap 0 = none)
    i64* work; i64 work_cap;        // (tile, own_lo, own_hi) triples: chunks the warp kernel hands to the generic kernel
    int list_mode;                  // generic kernel: iterate over `work` instead of all tiles
};

// ---------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D TMA bulk copy (SASS: UBLKCP / SYNCS)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// The corpus is read exactly once: its lines are marked evict-first in the L2, so that the stream (GBs) does not push the
// count table's hot sectors (tens of MB, probed over and over) out of the 126 MB L2.  PT_TEXT_EVICT_FIRST=0 restores the default.
#ifndef PT_TEXT_EVICT_FIRST
#define PT_TEXT_EVICT_FIRST 1
#endif
#ifndef PT_PROBE_EVICT_LAST
#define PT_PROBE_EVICT_LAST 0
#endif
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol)); return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol)); return pol;
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
#if PT_TEXT_EVICT_FIRST
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(l2_policy_evict_first()) : "memory");
#else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
#endif
}
// 16-byte table probe through the L2 only (ld.global.cg), optionally asking the L2 to keep the sector (evict-last)
__device__ __forceinline__ ulonglong2 probe_ld16(const void* p) {
#if PT_PROBE_EVICT_LAST
    ulonglong2 v;
    asm volatile("ld.global.cg.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;" : "=l"(v.x), "=l"(v.y) : "l"(p), "l"(l2_policy_evict_last()) : "memory");
    return v;
#else
    return __ldcg((const ulonglong2*)p);
#endif
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// ---------------------------------------------------------------------------------
// K1a: special-token candidates -> bitmap
// ---------------------------------------------------------------------------------
__device__ __forceinline__ i64 logical_end_after(const PretokParams& P, i64 i) {
    return hard_end_after(P.cuts, P.n_cuts, P.n, i);
}

__global__ void __launch_bounds__(256) k_special_candidates(PretokParams P, i64 lo, i64 hi) {
    // 16 bytes per thread and step; a SWAR zero-byte test finds the (rare) first bytes of specials
    const i64 stride = (i64)gridDim.x * blockDim.x;
    const i64 c0 = lo >> 4, c1 = (hi + 15) >> 4;
    const bool generic = c_sp.n_first > 8;
    for (i64 c = c0 + (i64)blockIdx.x * blockDim.x + threadIdx.x; c < c1; c += stride) {
        const uint4 v = ((const uint4*)P.text)[c];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t hit = 0;                                   // bit 4j+k: byte k of word j starts like a special
        if (!generic) {
            for (int f = 0; f < c_sp.n_first; f++) {
                const uint32_t F = 0x01010101u * c_sp.first[f];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t z = w[j] ^ F;
                    uint32_t m = ~(((z & 0x7f7f7f7fu) + 0x7f7f7f7fu) | z) & 0x80808080u;     // exact zero-byte flags
                    while (m) { int b = __ffs(m) - 1; m &= m - 1; hit |= 1u << (4 * j + (b >> 3)); }
                }
            }
        } else hit = 0xffffu;
        while (hit) {
            const int k = __ffs(hit) - 1; hit &= hit - 1;
            const i64 i = (c << 4) + k;
            if (i < lo || i >= hi) continue;
            const uint8_t b = P.text[i];
            if (!((c_sp.first_byte_mask[b >> 3] >> (b & 7)) & 1)) continue;
            const i64 lim = logical_end_after(P, i);
            if (special_match(P.text, i, lim) >= 0) atomicOr(&P.cand[i >> 5], 1u << (i & 31));
        }
    }
}

__device__ __forceinline__ bool bit_at(const uint32_t* bm, i64 i) { return (bm[i >> 5] >> (i & 31)) & 1; }
// next set bit in (from, to], or -1
__device__ i64 next_bit(const uint32_t* bm, i64 from, i64 to) {
    for (i64 i = from + 1; i <= to;) {
        uint32_t w = bm[i >> 5] >> (i & 31);
        if (w) { i64 r = i + __ffs(w) - 1; return r <= to ? r : -1; }
        i = ((i >> 5) + 1) << 5;
    }
    return -1;
}
__device__ bool any_bit(const uint32_t* bm, i64 lo, i64 hi) {   // any set bit in [lo, hi]
    if (lo < 0) lo = 0;
    if (hi < lo) return false;
    return next_bit(bm, lo - 1, hi) >= 0;
}

// K1b: walk chains of close candidates and decide which are recognised.
//   trainer: candidate q is recognised iff q is a token start given the previous recognised
//            special end as a fresh-text fence
//   encode : leftmost, priority order, non-overlapping (tokenizer.py:97-102,171)
// Candidates are sparse (about one per 30 bitmap words on document-separated text) while resolving one costs
// dozens of dependent loads: a thread-per-word scan would run that slow path with one or two active lanes in
// every warp iteration.  Each warp therefore first gathers candidate positions into its own shared-memory list
// (ballot + prefix) and runs the slow path 32 candidates at a time with all lanes busy.
__device__ __forceinline__ int resolve_candidate(const PretokParams& P, i64 q, i64 ctx_lo, int D) {
    // chain head?  no candidate in [q-D, q) -- nor before ctx_lo, a text start the caller vouches for (a hard cut: specials
    // never straddle one, so nothing to its left can reach q)
    if (D > 0 && any_bit(P.cand, q - D > ctx_lo ? q - D : ctx_lo, q - 1)) return 0;      // resolved by the walk of its chain's head
    GlobalText G{P.text, P.n, P.cuts, P.n_cuts, nullptr, -1, P.mode};
    int n_rec = 0;
    i64 e = -1, cur = q;
    for (;;) {
        i64 lim = logical_end_after(P, cur);
        int s = special_match(P.text, cur, lim);
        int m = c_sp.offs[s + 1] - c_sp.offs[s];
        bool ok = false;
        if (cur >= e) {
            if (P.mode == 1) ok = true;
            else { G.fence_fl = e; ok = is_token_start(G, cur); }
        }
        if (ok) { atomicOr(&P.rec[cur >> 5], 1u << (cur & 31)); e = cur + m; n_rec++; }
        i64 nx = D > 0 ? next_bit(P.cand, cur, cur + D < P.n - 1 ? cur + D : P.n - 1) : -1;
        if (nx < 0) break;
        cur = nx;
    }
    return n_rec;
}

__global__ void __launch_bounds__(256) k_resolve_specials(PretokParams P, i64 lo, i64 hi, i64 ctx_lo) {
    __shared__ i64 sh_list[8][64];
    const int D = P.mode == 0 ? c_sp.max_len + 16 : c_sp.max_len - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    i64* list = sh_list[warp];
    int cnt = 0;                                     // warp-uniform
    int my_rec = 0;
    const i64 nwords_lo = lo >> 5, nwords_hi = (hi + 31) >> 5;
    const i64 stride = (i64)gridDim.x * 8 * 32;
    for (i64 base = nwords_lo + ((i64)blockIdx.x * 8 + warp) * 32; base < nwords_hi; base += stride) {
        const i64 wi = base + lane;
        uint32_t w = wi < nwords_hi ? P.cand[wi] : 0u;
        while (__any_sync(0xffffffffu, w != 0)) {
            bool has = w != 0;
            i64 q = 0;
            if (has) { const int bit = __ffs(w) - 1; w &= w - 1; q = (wi << 5) + bit; has = q >= lo && q < hi; }
            const uint32_t m = __ballot_sync(0xffffffffu, has);
            if (has) list[cnt + __popc(m & lt)] = q;
            cnt += __popc(m);
            __syncwarp();
            if (cnt >= 32) {
                my_rec += resolve_candidate(P, list[lane], ctx_lo, D);
                __syncwarp();
                const i64 t = lane + 32 < cnt ? list[lane + 32] : 0;
                __syncwarp();
                if (lane + 32 < cnt) list[lane] = t;
                cnt -= 32;
                __syncwarp();
            }
        }
    }
    if (lane < cnt) my_rec += resolve_candidate(P, list[lane], ctx_lo, D);
    for (int o = 16; o > 0; o >>= 1) my_rec += __shfl_xor_sync(0xffffffffu, my_rec, o);
    if (lane == 0 && my_rec) atomicAdd((u64*)&P.stats[ST_NSPECIAL], (u64)my_rec);
}

// ---------------------------------------------------------------------------------
// hash-table inserts
// ---------------------------------------------------------------------------------
// multiply-add of the four 32-bit key words + one multiply-xorshift round: 9 integer instructions (the 64-bit
// mix it replaces cost ~60).  Table slot = low bits, shared-memory cache index = high bits.
__device__ __forceinline__ uint32_t short_hash_w(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    uint32_t x = a * 0x9E3779B1u + b * 0x85EBCA77u + c * 0xC2B2AE3Du + d * 0x27D4EB2Fu;
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15;
    return x;
}
__device__ __forceinline__ u64 short_hash(u64 k0, u64 k1) {
    return short_hash_w((uint32_t)k0, (uint32_t)(k0 >> 32), (uint32_t)k1, (uint32_t)(k1 >> 32));
}

// returns slot (>=0) and adds `add` to its count; *created = 1 when this call created the entry
// 16-byte compare-and-swap (atom.cas.b128, sm_90+): a slot is claimed with ONE atomic instead of a word-by-word protocol
__device__ __forceinline__ ulonglong2 atom_cas128(void* p, u64 c0, u64 c1, u64 v0, u64 v1) {
    ulonglong2 old;
    asm volatile("{\n\t.reg .b128 c, v, o;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 v, {%4, %5};\n\t"
                 "atom.global.cas.b128 o, [%6], c, v;\n\tmov.b128 {%0, %1}, o;\n\t}"
                 : "=l"(old.x), "=l"(old.y) : "l"(c0), "l"(c1), "l"(v0), "l"(v1), "l"(p) : "memory");
    return old;
}

// `seen` = what a previous 16-byte load of the first slot returned (saves the first read), or {~0, ~0} for "not read yet"
__device__ __forceinline__ i64 short_insert_seen(const ShortTab& T, u64 h, u64 k0, u64 k1, i64 add, int* created, ulonglong2 seen) {
    u64 mask = (u64)T.cap - 1;
    u64 slot = h & mask;
    *created = 0;
#pragma unroll 1
    for (int probe = 0; probe < 8192; probe++) {
        ulonglong2 kv = seen;
        if (kv.x == ~0ULL) kv = probe_ld16(T.key(slot));
        seen.x = ~0ULL;
        if (kv.x == 0 && kv.y == 0) {
            kv = atom_cas128(T.key(slot), 0ULL, 0ULL, k0, k1);
            if (kv.x == 0 && kv.y == 0) { *created = 1; atomicAdd((u64*)T.cnt(slot), (u64)add); return (i64)slot; }
        }
        if (kv.x == k0 && kv.y == k1) { atomicAdd((u64*)T.cnt(slot), (u64)add); return (i64)slot; }
        if ((kv.x == 0) != (kv.y == 0)) continue;          // half a key (both words are non-zero in every key): a torn read, look again
        slot = (slot + 1) & mask;
    }
    return -1;
}
__device__ __forceinline__ i64 short_insert_h(const ShortTab& T, u64 h, u64 k0, u64 k1, i64 add, int* created) {
    ulonglong2 none; none.x = ~0ULL; none.y = ~0ULL;
    return short_insert_seen(T, h, k0, k1, add, created, none);
}

__device__ __forceinline__ i64 short_insert(const ShortTab& T, u64 k0, u64 k1, i64 add, int* created) {
    return short_insert_h(T, short_hash(k0, k1), k0, k1, add, created);
}

// Per-CTA pre-aggregation cache in shared memory: hot pre-tokens (Zipf head) are counted with
// shared-memory atomics and reach the global table once per CTA, at the end of the kernel.
#ifndef PT_CACHE_N
#define PT_CACHE_N 1024
#endif
#define PT_CACHE_PROBES 4
#define PT_CACHE_BYTES (PT_CACHE_N * 20)

__device__ __forceinline__ bool cache_add(u64* ck0, u64* ck1, uint32_t* cc, u64 h, u64 k0, u64 k1) {
    uint32_t ci = (uint32_t)(h >> 16) & (PT_CACHE_N - 1);
#pragma unroll
    for (int p = 0; p < PT_CACHE_PROBES; p++) {
        u64 c0 = *(volatile u64*)&ck0[ci];
        if (c0 == 0) { c0 = atomicCAS(&ck0[ci], 0ULL, k0); if (c0 == 0) c0 = k0; }
        if (c0 == k0) {
            u64 c1 = *(volatile u64*)&ck1[ci];
            if (c1 == 0) { c1 = atomicCAS(&ck1[ci], 0ULL, k1); if (c1 == 0) c1 = k1; }
            if (c1 == k1) { atomicAdd(&cc[ci], 1u); return true; }
        }
        ci = (ci + 1) & (PT_CACHE_N - 1);
    }
    return false;
}

// read-only lookup (encode passes); -1 when absent
__device__ __forceinline__ i64 short_find(const ShortTab& T, u64 k0, u64 k1) {
    u64 mask = (u64)T.cap - 1;
    u64 slot = short_hash(k0, k1) & mask;
    for (int probe = 0; probe < 8192; probe++) {
        const ulonglong2 kv = *(const ulonglong2*)T.key(slot);
        if (kv.x == k0 && kv.y == k1) return (i64)slot;
        if (kv.x == 0) return -1;
        slot = (slot + 1) & mask;
    }
    return -1;
}

// lookup of the per-slot record (encode passes after k_encode_finalize): key and record sit in the same sector
__device__ __forceinline__ i64 short_find_info(const ShortTab& T, u64 k0, u64 k1) {
    u64 mask = (u64)T.cap - 1;
    u64 slot = short_hash(k0, k1) & mask;
    for (int probe = 0; probe < 8192; probe++) {
        const ulonglong2 kv = *(const ulonglong2*)T.key(slot);
        const i64 info = *T.cnt(slot);
        if (kv.x == k0 && kv.y == k1) return info;
        if (kv.x == 0) return -1;
        slot = (slot + 1) & mask;
    }
    return -1;
}

__device__ __forceinline__ u64 long_hash_fix(u64 h) { return h < 2 ? h + 2 : h; }

__device__ bool text_equal(const uint8_t* text, i64 a, i64 b, i64 len) {
    for (i64 k = 0; k < len; k++) if (text[a + k] != text[b + k]) return false;
    return true;
}

// thread-level insert of a long pre-token text[pos, pos+len) with hash h (exact: bytes compared)
__device__ i64 long_insert(LongEntry* ent, i64 cap, const uint8_t* text, u64 h, i64 pos, i64 len, i64 add, int* created) {
    u64 mask = (u64)cap - 1;
    u64 slot = h & mask;
    *created = 0;
    int probes = 0;
    for (;;) {
        u64* hp = &ent[slot].h;
        u64 cur = *(volatile u64*)hp;
        if (cur == 0) {
            cur = atomicCAS(hp, 0ULL, 1ULL);
            if (cur == 0) {
                ent[slot].pos = pos; ent[slot].len = len; ent[slot].count = 0;
                __threadfence();
                atomicExch(hp, h);
                atomicAdd((u64*)&ent[slot].count, (u64)add);
                *created = 1;
                return (i64)slot;
            }
        }
        if (cur == 1) continue;                       // being published by another thread: retry this slot
        if (cur == h) {
            __threadfence();
            i64 elen = *(volatile i64*)&ent[slot].len, epos = *(volatile i64*)&ent[slot].pos;
            if (elen == len && (epos == pos || text_equal(text, epos, pos, len))) {
                atomicAdd((u64*)&ent[slot].count, (u64)add);
                return (i64)slot;
            }
        }
        slot = (slot + 1) & mask;
        if (++probes > 8192) return -1;
    }
}

__device__ i64 long_find(const LongEntry* ent, i64 cap, const uint8_t* text, u64 h, i64 pos, i64 len) {
    u64 mask = (u64)cap - 1;
    u64 slot = h & mask;
    for (int probes = 0; probes < 8192; probes++) {
        u64 cur = ent[slot].h;
        if (cur == 0) return -1;
        if (cur == h && ent[slot].len == len && (ent[slot].pos == pos || text_equal(text, ent[slot].pos, pos, len))) return (i64)slot;
        slot = (slot + 1) & mask;
    }
    return -1;
}

// ---------------------------------------------------------------------------------
// tile machinery
// ---------------------------------------------------------------------------------
struct TileSmem {
    alignas(128) uint8_t txt[2][PT_WIN];
    union {                                   // info: generic passes A-C only; tokpos: from pass D1 on
        alignas(16) uint8_t info[PT_WIN];
        uint16_t tokpos[PT_TILE + 8];
    };
    uint32_t smask[PT_NMASK + 7];
    uint8_t lut[256];
    alignas(8) uint64_t bar[2];
    int scan_tmp[PT_THREADS / 32];
    int scan_carry;
    int ntok_total, ntok_own;
    // fast (pure-ASCII interior tile) path
    uint32_t recw[PT_WIN / 32 + 8];      // recognised-special bits; word i covers window offsets [32(i-4), 32(i-3))
};

__device__ __forceinline__ int block_exclusive_scan(int v, int* tmp, int* total) {
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) tmp[wid] = inc;
    __syncthreads();
    int base = 0, tot = 0;
    for (int k = 0; k < PT_THREADS / 32; k++) { int t = tmp[k]; if (k < wid) base += t; tot += t; }
    __syncthreads();
    *total = tot;
    return base + inc - v;
}

// issue the TMA load of tile `t` into buffer `buf`; bytes outside [0, n16) are zero-filled by the caller
__device__ __forceinline__ void tile_issue_load(const PretokParams& P, TileSmem& S, i64 tile, int buf) {
    i64 t0 = (P.tile_base + tile) * PT_TILE;
    i64 g0 = t0 - PT_HL, g1 = g0 + PT_WIN;
    i64 n16 = (P.n + 15) & ~(i64)15;
    i64 lo = g0 < 0 ? 0 : g0, hi = g1 > n16 ? n16 : g1;
    if (hi > lo) {
        uint32_t bytes = (uint32_t)(hi - lo);
        mbar_expect_tx(&S.bar[buf], bytes);
        tma_load_1d(&S.txt[buf][lo - g0], P.text + lo, bytes, &S.bar[buf]);
    } else {
        mbar_expect_tx(&S.bar[buf], 0);
    }
}

__device__ __forceinline__ void info_or(uint8_t* info, int x, uint32_t bits) {
    atomicOr((uint32_t*)(info + (x & ~3)), bits << ((x & 3) * 8));
}

// Generic path, passes A-C: any text (multi-byte code points, hard cuts, buffer ends).  Writes S.smask.
__device__ void tile_scan_generic(const PretokParams& P, TileSmem& S, i64 tile, int buf) {
    const int tid = threadIdx.x;
    const i64 t0 = (P.tile_base + tile) * PT_TILE;
    const i64 g0 = t0 - PT_HL;
    uint8_t* txt = S.txt[buf];
    uint8_t* info = S.info;
    const i64 n16 = (P.n + 15) & ~(i64)15;

    // zero the parts of the window that TMA did not fill, and bytes in [n, n16)
    {
        i64 lo = g0 < 0 ? 0 : g0, hi = g0 + PT_WIN > n16 ? n16 : g0 + PT_WIN;
        for (int x = tid; x < PT_WIN; x += PT_THREADS) {
            i64 g = g0 + x;
            if (g < lo || g >= hi || g >= P.n) txt[x] = 0;
        }
    }
    __syncthreads();

    // ---- pass A1: base info from the byte LUT
    for (int k = tid; k < PT_WIN / 4; k += PT_THREADS) {
        uint32_t w = ((const uint32_t*)txt)[k];
        uint32_t v = (uint32_t)S.lut[w & 0xff] | ((uint32_t)S.lut[(w >> 8) & 0xff] << 8) |
                     ((uint32_t)S.lut[(w >> 16) & 0xff] << 16) | ((uint32_t)S.lut[w >> 24] << 24);
        i64 g = g0 + 4 * (i64)k;
        if (g + 3 >= P.n || g <= 0) {     // logical ends of the buffer
            for (int j = 0; j < 4; j++) {
                i64 gj = g + j;
                if (gj >= P.n || gj < 0) v = (v & ~(0xffu << (8 * j))) | ((uint32_t)(IB_FL | IB_FR) << (8 * j));
                else if (gj == 0) v |= (uint32_t)IB_FL << (8 * j);
            }
        }
        ((uint32_t*)info)[k] = v;
    }
    __syncthreads();

    // ---- pass A3: fences from chunk cuts and recognised specials
    if (P.n_cuts > 0) {
        int lo = 0, hi = P.n_cuts;
        i64 first = g0 < 1 ? 1 : g0;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (P.cuts[mid] < first) lo = mid + 1; else hi = mid; }
        for (int c = lo + tid; c < P.n_cuts; c += PT_THREADS) {
            i64 cp = P.cuts[c];
            if (cp >= g0 + PT_WIN) break;
            if (cp < P.n) info_or(info, (int)(cp - g0), IB_FL | IB_FR);
        }
    }
    if (P.n_sp > 0) {
        i64 lo = g0 - c_sp.max_len; if (lo < 0) lo = 0;
        i64 hi = g0 + PT_WIN; if (hi > P.n) hi = P.n;
        for (i64 wi = (lo >> 5) + tid; wi <= ((hi - 1) >> 5) && hi > 0; wi += PT_THREADS) {
            uint32_t w = P.rec[wi];
            while (w) {
                int bit = __ffs(w) - 1; w &= w - 1;
                i64 q = (wi << 5) + bit;
                if (q < lo || q >= hi) continue;
                int s = special_match(P.text, q, logical_end_after(P, q));
                if (s < 0) continue;
                int m = c_sp.offs[s + 1] - c_sp.offs[s];
                for (int k = 0; k <= m; k++) {
                    i64 x = q + k - g0;
                    if (x < 0 || x >= PT_WIN) continue;
                    uint32_t bits = k < m ? IB_IN : 0;
                    if (k == 0) bits |= IB_FL | (P.mode == 1 ? IB_FR : 0);
                    if (k == m) bits |= IB_FL;
                    info_or(info, (int)x, bits);
                }
            }
        }
    }
    __syncthreads();

    // ---- pass A2: multi-byte code points (class propagated to continuation bytes) + strict UTF-8
    for (int k = tid; k < PT_WIN / 4; k += PT_THREADS) {
        uint32_t w = ((const uint32_t*)txt)[k];
        if (!(w & 0x80808080u)) continue;
        for (int j = 0; j < 4; j++) {
            int x = 4 * k + j;
            uint8_t b = txt[x];
            if (b < 0x80) continue;
            i64 g = g0 + x;
            if (g >= P.n) continue;
            bool own = g >= t0 && g < t0 + PT_TILE;
            if ((b & 0xC0) == 0x80) {
                if (!own) continue;
                // stray continuation byte?  must be covered by a lead within 3 bytes
                bool covered = false;
                for (int d = 1; d <= 3 && x - d >= 0; d++) {
                    uint8_t pb = txt[x - d];
                    if (info[x - d + 1] & IB_FL) break;        // a fence between lead and this byte
                    if ((pb & 0xC0) == 0x80) continue;
                    covered = pb >= 0xC0 && utf8_len_from_lead(pb) > d;
                    break;
                }
                if (!covered) atomicMin((i64*)&P.stats[ST_ERR_POS], g);
                continue;
            }
            int len; bool ok = x + 4 <= PT_WIN ? utf8_seq_ok(txt + x, P.n - g, &len) : false;
            if (x + 4 > PT_WIN) continue;       // slack region: never needed as a class
            if (ok) for (int d = 1; d < len; d++) if (info[x + d] & IB_FR) ok = false;
            if (!ok) { if (own) atomicMin((i64*)&P.stats[ST_ERR_POS], g); continue; }
            int kc = kclass_of_cp(utf8_decode(txt + x, len));
            if (kc) for (int d = 0; d < len; d++) info_or(info, x + d, (uint32_t)kc);
        }
    }
    __syncthreads();

    // ---- pass B: live contractions
    for (int k = tid; k < PT_WIN / 4; k += PT_THREADS) {
        uint32_t w = ((const uint32_t*)txt)[k] ^ 0x27272727u;
        if (!((w - 0x01010101u) & ~w & 0x80808080u)) continue;
        for (int j = 0; j < 4; j++) {
            int a = 4 * k + j;
            if (txt[a] != '\'' || a + 3 >= PT_WIN || a < 1) continue;
            uint8_t ia = info[a];
            if (ia & IB_IN) continue;
            if (info[a + 1] & IB_FR) continue;
            uint8_t c1 = txt[a + 1], c2 = txt[a + 2];
            int clen = 0;
            if (c1 == 's' || c1 == 'd' || c1 == 'm' || c1 == 't') clen = 2;
            else if (!(info[a + 2] & IB_FR) && ((c1 == 'l' && c2 == 'l') || (c1 == 'v' && c2 == 'e') || (c1 == 'r' && c2 == 'e'))) clen = 3;
            if (!clen) continue;
            bool live = ia & IB_FL;
            if (!live) { int p = info[a - 1] & IB_CLS; live = (p == KC_L || p == KC_N || p == KC_S); }
            if (!live) continue;
            for (int d = 1; d < clen; d++) info_or(info, a + d, IB_INC);
            info_or(info, a + clen, IB_FL);      // the position after a contraction behaves like a text start
        }
    }
    __syncthreads();

    // ---- pass C: start bit per byte, one ballot per 32 bytes
    for (int base = 0; base < PT_NMASK * 32; base += PT_THREADS) {
        int b = base + tid;              // bit index; window offset x = HL + b
        int x = PT_HL + b;
        bool start = false;
        if (b < PT_NBITS) {
            uint8_t v = info[x];
            if (v & IB_CONT) start = false;
            else if (v & IB_FL) start = true;
            else if (v & (IB_IN | IB_INC)) start = false;
            else {
                int c = v & IB_CLS, p = info[x - 1] & IB_CLS;
                if (c < KC_S) start = (p == KC_SP) ? false : (p >= KC_S ? true : p != c);
                else if (p < KC_S) start = true;
                else {
                    uint8_t nv = info[x + utf8_len_from_lead(txt[x])];
                    start = !(nv & IB_FR) && (nv & IB_CLS) < KC_S;
                }
            }
        }
        uint32_t m = __ballot_sync(0xffffffffu, start);
        if ((tid & 31) == 0 && (b >> 5) < PT_NMASK) S.smask[b >> 5] = m;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------
// Fast path: interior tile, pure ASCII.  One thread scans one 32-byte segment entirely in
// registers: SWAR classification of the 8 words, predicate bits gathered into "transposed"
// 32-bit masks (bit 8k+j = byte k of word j), the start rule as ~25 bitwise operations, and
// rare per-thread events (recognised specials, live contractions) patched in afterwards.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ bool rec_bit_g(const PretokParams& P, i64 q) { return P.n_sp > 0 && ((P.rec[q >> 5] >> (q & 31)) & 1); }
__device__ __forceinline__ uint32_t tbit(int p) { return 1u << (((p & 3) << 3) | (p >> 2)); }       // segment position -> mask bit
__device__ __forceinline__ uint32_t prevT(uint32_t m, uint32_t carry) { return (m << 8) | (((m >> 24) << 1) & 0xffu) | carry; }
__device__ __forceinline__ uint32_t nextT(uint32_t m, uint32_t carry) { return (m >> 8) | (((m & 0xffu) >> 1) << 24) | (carry << 31); }
__device__ __forceinline__ uint32_t spread8(uint32_t b) {     // bit j -> bit 4j
    b = (b | (b << 12)) & 0x000f000fu;
    b = (b | (b << 6)) & 0x03030303u;
    b = (b | (b << 3)) & 0x11111111u;
    return b;
}
__device__ __forceinline__ uint32_t untranspose(uint32_t t) {
    return spread8(t & 0xff) | (spread8((t >> 8) & 0xff) << 1) | (spread8((t >> 16) & 0xff) << 2) | (spread8(t >> 24) << 3);
}
__device__ __forceinline__ int ascii_kclass(uint32_t b) {
    if (((b | 0x20) - 'a') < 26u) return KC_L;
    if ((b - '0') < 10u) return KC_N;
    if (b == 0x20) return KC_SP;
    if ((b - 9) < 5u) return KC_S;
    return KC_O;
}

// class of the code point that window byte `pos` belongs to (multi-byte aware; O when ill-formed)
__device__ int smem_kclass(const uint8_t* txt, int pos) {
    uint8_t b = txt[pos];
    if (b < 0x80) return ascii_kclass(b);
    int y = pos;
    while ((txt[y] & 0xC0) == 0x80 && pos - y < 3) y--;
    b = txt[y];
    if (b < 0xC0) return KC_O;
    int len;
    if (!utf8_seq_ok(txt + y, 8, &len) || y + len <= pos) return KC_O;
    return kclass_of_cp(utf8_decode(txt + y, len));
}

// window touches offset 0 / n or contains a hard cut -> generic path (every thread computes the same answer)
__device__ __forceinline__ bool tile_is_boundary(const PretokParams& P, i64 g0, bool* has_cut) {
    bool cut = false;
    if (P.n_cuts > 0) {
        int lo = 0, hi = P.n_cuts;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (P.cuts[mid] < g0) lo = mid + 1; else hi = mid; }
        cut = lo < P.n_cuts && P.cuts[lo] <= g0 + PT_WIN;
    }
    *has_cut = cut;
    return cut || g0 <= 0 || g0 + PT_WIN >= P.n;
}

// stage the recognised-special bits of the window (segment_scan; encode mode: also the token loops after pass D1)
__device__ __forceinline__ void tile_stage_rec(const PretokParams& P, TileSmem& S, i64 g0) {
    if (P.n_sp == 0) return;
    const i64 nrec = (P.n + 63) / 32 + 1;
    for (int i = threadIdx.x; i < PT_WIN / 32 + 8; i += PT_THREADS) {
        i64 wi = (g0 >> 5) - 4 + i;              // g0 is a multiple of 32
        S.recw[i] = (wi >= 0 && wi < nrec) ? P.rec[wi] : 0u;
    }
}

// scan one 32-byte segment (index sgi, window offset x0 = HL + 32*sgi); returns the start mask in natural
// bit order.  recw = recognised-special bits of the window staged in shared memory: word (x >> 5) + 4 covers
// window offsets [x & ~31, (x & ~31) + 32) (the window base is a multiple of 32; four words of history).
// Class masks are "transposed" (tbit); the rare-event masks FL / SUP / KILL are in natural bit order so that
// a special or a contraction patches a bit RANGE with two shifts instead of a per-byte loop.
__device__ __forceinline__ uint32_t range_mask(int lo, int hi) {        // bits [lo, hi), 0 <= lo < hi <= 32
    return (hi - lo >= 32 ? 0xffffffffu : ((1u << (hi - lo)) - 1u)) << lo;
}
__device__ __forceinline__ int contraction_len(uint8_t c1, uint8_t c2) {
    if (c1 == 's' || c1 == 'd' || c1 == 'm' || c1 == 't') return 2;
    if ((c1 == 'l' && c2 == 'l') || (c1 == 'v' && c2 == 'e') || (c1 == 'r' && c2 == 'e')) return 3;
    return 0;
}
__device__ __forceinline__ uint32_t segment_scan(const PretokParams& P, const uint8_t* txt, const uint32_t* recw, i64 g0, int sgi) {
    const int x0 = PT_HL + 32 * sgi;
    const uint4* p4 = (const uint4*)(txt + x0);
    const uint4 A = p4[0], B = p4[1];
    const uint32_t w[8] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w};
    const uint32_t pw = ((const uint32_t*)txt)[x0 / 4 - 1], nw = ((const uint32_t*)txt)[x0 / 4 + 8];
    uint32_t na = (pw | nw) & 0x80808080u;
    uint32_t Lm = 0, Nm = 0, Sm = 0, SPm = 0, APm = 0, NAm = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        na |= w[j] & 0x80808080u;
        NAm |= (w[j] & 0x80808080u) >> (7 - j);
        const uint32_t x = w[j] & 0x7f7f7f7fu, y = x | 0x20202020u;
        const uint32_t fL = (y + 0x1f1f1f1fu) & ~(y + 0x05050505u) & 0x80808080u;
        const uint32_t fN = (x + 0x50505050u) & ~(x + 0x46464646u) & 0x80808080u;
        const uint32_t fSP = ~((x ^ 0x20202020u) + 0x7f7f7f7fu) & 0x80808080u;
        const uint32_t fS = ((x + 0x77777777u) & ~(x + 0x72727272u) & 0x80808080u) | fSP;
        const uint32_t fAP = ~((x ^ 0x27272727u) + 0x7f7f7f7fu) & 0x80808080u;
        Lm |= fL >> (7 - j); Nm |= fN >> (7 - j); Sm |= fS >> (7 - j); SPm |= fSP >> (7 - j); APm |= fAP >> (7 - j);
    }
    const uint32_t pb = pw >> 24, nb = nw & 0xffu;
    int pk = ascii_kclass(pb);
    uint32_t nS_in = (nb == 0x20 || (nb - 9) < 5u) ? 1u : 0u;
    uint32_t FLn = 0, SUPn = 0, KILLn = 0;          // natural order: forced starts, suppressed starts, bytes of specials
    uint32_t NOEXT = 0, CONT = 0, wsleads = 0;      // transposed order
    // ---- events: non-ASCII code points (decoded one by one; classes from the two-stage Unicode table)
    if (na) {
        Lm &= ~NAm; Nm &= ~NAm; Sm &= ~NAm; SPm &= ~NAm; APm &= ~NAm;     // the SWAR tests looked at the low 7 bits only
        uint32_t covered = 0;
        uint32_t nat = untranspose(NAm);
        for (int it = -3; it < 32; it++) {
            if (it >= 0) { if (!nat) break; it = __ffs(nat) - 1; nat &= nat - 1; }
            const int x = x0 + it;
            const uint8_t b = txt[x];
            if (b < 0xC0) continue;                                   // ASCII or continuation byte
            int len;
            const bool ok = utf8_seq_ok(txt + x, P.n - (g0 + x), &len);
            if (it < 0 && len <= -it) continue;                       // ends before the segment
            if (!ok) { if (it >= 0) atomicMin((i64*)&P.stats[ST_ERR_POS], g0 + x); continue; }
            const int kc = kclass_of_cp(utf8_decode(txt + x, len));
            for (int t = 0; t < len; t++) {
                const int pos = it + t;
                if (pos < 0 || pos >= 32) continue;
                const uint32_t tb = tbit(pos);
                covered |= tb;
                if (t > 0) CONT |= tb;
                if (kc == KC_L) Lm |= tb; else if (kc == KC_N) Nm |= tb; else if (kc == KC_S) Sm |= tb;
            }
            if (kc == KC_S && it >= 0) wsleads |= tbit(it);
        }
        uint32_t stray = untranspose(NAm & ~covered);                 // bytes >= 0x80 not part of a well-formed sequence
        if (stray) atomicMin((i64*)&P.stats[ST_ERR_POS], g0 + x0 + __ffs(stray) - 1);
        if (pb >= 0x80) pk = smem_kclass(txt, x0 - 1);
        if (nb >= 0x80) nS_in = smem_kclass(txt, x0 + 32) >= KC_S ? 1u : 0u;
    }
    uint32_t flprev = 0, killprev = 0;           // bit d-1: position x0-d (d = 1..3) is a forced start / inside a special
    // ---- events: recognised specials that touch [x0-3, x0+32]
    if (P.n_sp > 0) {
        const int wi = (x0 >> 5) + 4;
        const int kprev = (c_sp.max_len + 3 + 31) >> 5;              // <= 5 (specials are at most 128 bytes)
        uint32_t any = recw[wi] | (recw[wi + 1] & 1u) | recw[wi - 1];
#pragma unroll 1
        for (int j = 2; j <= kprev; j++) any |= recw[wi - j];
        if (any) {
#pragma unroll 1
            for (int j = -kprev; j <= 1; j++) {
                uint32_t rw = recw[wi + j];
                if (j == 1) rw &= 1u;                                // only a special starting right after the segment matters
                while (rw) {
                    const int bit = __ffs(rw) - 1; rw &= rw - 1;
                    const int q = x0 + 32 * j + bit;                 // window offset
                    int m;
                    if (c_sp.n == 1) m = c_sp.offs[1];
                    else { int sp = special_match(P.text, g0 + q, logical_end_after(P, g0 + q)); m = sp >= 0 ? c_sp.offs[sp + 1] - c_sp.offs[sp] : 1; }
                    const int e = q + m;
                    if (e < x0 - 3 || q > x0 + 32) continue;
                    const int lo = (q > x0 ? q : x0) - x0, hi = (e < x0 + 32 ? e : x0 + 32) - x0;
                    if (hi > lo) {
                        const uint32_t r = range_mask(lo, hi);
                        KILLn |= r;
                        SUPn |= q >= x0 ? (r & ~(1u << lo)) : r;
                    }
#pragma unroll
                    for (int d = 1; d <= 3; d++) if (q <= x0 - d && x0 - d < e) killprev |= 1u << (d - 1);
                    if (q >= x0 && q < x0 + 32) FLn |= 1u << (q - x0);
                    if (e >= x0 && e < x0 + 32) FLn |= 1u << (e - x0);
                    else if (e >= x0 - 3 && e < x0) flprev |= 1u << (x0 - e - 1);
                    if (P.mode == 1) {                  // encode: the text before a special ends at it
                        if (q - 1 >= x0 && q - 1 < x0 + 32) NOEXT |= tbit(q - 1 - x0);
                        if (q == x0 + 32) nS_in = 1u;
                    }
                }
            }
        }
    }
    const uint32_t pL = prevT(Lm, pk == KC_L), pN = prevT(Nm, pk == KC_N), pS = prevT(Sm, pk >= KC_S), pSP = prevT(SPm, pk == KC_SP);
    // ---- events: live contractions whose effects reach this segment (apostrophe at x0-3 .. x0+31)
    uint32_t apprev = 0;
    if (((pw >> 8) & 0xff) == '\'') apprev |= 4u;      // x0-3
    if (((pw >> 16) & 0xff) == '\'') apprev |= 2u;     // x0-2
    if ((pw >> 24) == '\'') apprev |= 1u;              // x0-1
    apprev &= ~killprev;
    if (APm | apprev) {
        uint32_t nat = untranspose(APm) & ~KILLn;
        for (int d = 3; d >= 1; d--) {
            if (!(apprev & (1u << (d - 1)))) continue;
            const int a = x0 - d;
            int clen = contraction_len(txt[a + 1], txt[a + 2]);
            if (P.mode == 1 && clen) {          // the contraction may not reach into a special (text ends there)
                if (rec_bit_g(P, g0 + a + 1)) clen = 0;
                else if (clen == 3 && (rec_bit_g(P, g0 + a + 2))) clen = 0;
            }
            if (clen < d) continue;              // ends before this segment
            bool live = (flprev >> (d - 1)) & 1;
            if (!live) { int k = smem_kclass(txt, a - 1); live = (k == KC_L || k == KC_N || k == KC_S); }
            if (!live) continue;
            if (clen > d) SUPn |= range_mask(0, clen - d);     // positions a+1 .. a+clen-1 that fall into the segment
            if (a + clen < x0 + 32) FLn |= 1u << (a + clen - x0);   // the position after a contraction is a forced start
        }
        while (nat) {
            const int ap = __ffs(nat) - 1; nat &= nat - 1;
            const int a = x0 + ap;
            const int clen = contraction_len(txt[a + 1], txt[a + 2]);
            if (!clen) continue;
            if (P.mode == 1) {
                if (rec_bit_g(P, g0 + a + 1)) continue;
                if (clen == 3 && (rec_bit_g(P, g0 + a + 2))) continue;
            }
            const uint32_t tb = tbit(ap);
            const bool live = ((FLn >> ap) & 1u) || (pL & tb) || (pN & tb) || ((pS & tb) && !(pSP & tb));
            if (!live) continue;
            const int hi = ap + clen < 32 ? ap + clen : 32;
            if (hi > ap + 1) SUPn |= range_mask(ap + 1, hi);
            if (ap + clen < 32) FLn |= 1u << (ap + clen);
        }
    }
    // ---- the start rule, 32 positions at once
    const uint32_t Om = ~(Lm | Nm | Sm), pO = ~(pL | pN | pS);
    const uint32_t same = (Lm & pL) | (Nm & pN) | (Om & pO);
    const uint32_t NS = nextT(Sm, nS_in) | NOEXT;
    uint32_t st = (untranspose(((~Sm & ~pSP & (pS | ~same)) | (Sm & (~pS | ~NS))) & ~CONT) & ~SUPn) | FLn;
    // multi-byte whitespace (U+00A0, U+2003, U+3000, ...): the look-ahead is the next CODE POINT, not the next byte
    if (wsleads) {
        uint32_t nat = untranspose(wsleads);
        while (nat) {
            const int pos = __ffs(nat) - 1; nat &= nat - 1;
            if (((FLn | SUPn) >> pos) & 1u) continue;                // already decided
            if (!(pS & tbit(pos))) continue;                         // previous is not whitespace: a start, as computed
            const int nx = x0 + pos + utf8_len_from_lead(txt[x0 + pos]);
            const bool exists = !(P.mode == 1 && rec_bit_g(P, g0 + nx));
            const bool start = exists && smem_kclass(txt, nx) < KC_S;
            st = start ? (st | (1u << pos)) : (st & ~(1u << pos));
        }
    }
    return st;
}

// Fast path, whole tile.  Returns false when the tile needs the generic path (smask is then rewritten).
__device__ bool tile_scan_fast(const PretokParams& P, TileSmem& S, i64 g0, int buf) {
    const uint8_t* txt = S.txt[buf];
    for (int sgi = threadIdx.x; sgi < PT_NSEG; sgi += PT_THREADS) {
        uint32_t m = segment_scan(P, txt, S.recw, g0, sgi);
        S.smask[sgi] = sgi == PT_NSEG - 1 ? (m & 1u) : m;
    }
    __syncthreads();
    return true;
}

// pass D1: compact the start positions of S.smask into S.tokpos.  Only starts inside the tile become
// tokens; the first start in the right halo is kept as the end of the last one.
__device__ void tile_compact(TileSmem& S) {
    const int tid = threadIdx.x;
    int carry = 0;
    for (int base = 0; base < PT_TILE / 32; base += PT_THREADS) {
        int wi = base + tid;
        uint32_t m = wi < PT_TILE / 32 ? S.smask[wi] : 0;
        int cnt = __popc(m), total;
        int off = block_exclusive_scan(cnt, S.scan_tmp, &total) + carry;
        while (m) { int bit = __ffs(m) - 1; m &= m - 1; S.tokpos[off++] = (uint16_t)(PT_HL + wi * 32 + bit); }
        carry += total;
    }
    if (tid == 0) {
        int sentinel = -1;
        for (int wi = PT_TILE / 32; wi < PT_NMASK && sentinel < 0; wi++) { uint32_t m = S.smask[wi]; if (m) sentinel = PT_HL + wi * 32 + __ffs(m) - 1; }
        S.ntok_own = carry;
        S.ntok_total = carry + (sentinel >= 0 ? 1 : 0);
        if (sentinel >= 0) S.tokpos[carry] = (uint16_t)sentinel;
    }
    __syncthreads();
}

// Passes A-D1 for one tile.  On return S.tokpos / S.ntok_* describe the tile's tokens.
__device__ void tile_scan(const PretokParams& P, TileSmem& S, i64 tile, int buf, bool* has_cut) {
    const i64 g0 = (P.tile_base + tile) * PT_TILE - PT_HL;
    tile_stage_rec(P, S, g0);
    if (P.n_sp > 0) __syncthreads();
    bool done = false;
    if (!tile_is_boundary(P, g0, has_cut)) done = tile_scan_fast(P, S, g0, buf);
    if (!done) tile_scan_generic(P, S, tile, buf);
    tile_compact(S);
}

// 16 bytes of the window starting at offset s (unaligned) as two u64
__device__ __forceinline__ void load16(const uint8_t* txt, int s, u64* lo, u64* hi) {
    const uint32_t* tw = (const uint32_t*)txt;
    int wi = s >> 2, sh = (s & 3) * 8;
    uint32_t a0 = tw[wi], a1 = tw[wi + 1], a2 = tw[wi + 2], a3 = tw[wi + 3], a4 = tw[wi + 4];
    uint32_t b0 = __funnelshift_r(a0, a1, sh), b1 = __funnelshift_r(a1, a2, sh);
    uint32_t b2 = __funnelshift_r(a2, a3, sh), b3 = __funnelshift_r(a3, a4, sh);
    *lo = (u64)b0 | ((u64)b1 << 32);
    *hi = (u64)b2 | ((u64)b3 << 32);
}

// pack a pre-token of len <= 14 into the two key words (len tag in the top byte of k0)
__device__ __forceinline__ void pack_short_key(const uint8_t* txt, int s, int len, u64* k0, u64* k1) {
    u64 lo, hi;
    load16(txt, s, &lo, &hi);
    if (len < 8) { lo &= (1ULL << (8 * len)) - 1; hi = 0; }
    else if (len < 16) hi &= (1ULL << (8 * (len - 8))) - 1;
    *k0 = (lo & 0x00FFFFFFFFFFFFFFULL) | ((u64)len << 56);
    *k1 = (lo >> 56) | (hi << 8) | (1ULL << 56);
}

__device__ __forceinline__ void init_tile_smem(TileSmem& S) {
    // byte LUT: ASCII classes, continuation flag; leads >= 0xC0 get class in pass A2
    for (int b = threadIdx.x; b < 256; b += PT_THREADS) {
        uint8_t v = 0;
        if (b < 0x80) v = (uint8_t)kclass_of_cp((uint32_t)b);
        else if (b < 0xC0) v = IB_CONT;
        S.lut[b] = v;
    }
    if (threadIdx.x == 0) { mbar_init(&S.bar[0], 1); mbar_init(&S.bar[1], 1); mbar_fence_init(); }
    __syncthreads();
}

// ---------------------------------------------------------------------------------
// K1+K2: count pre-tokens into the hash tables
// ---------------------------------------------------------------------------------
#ifndef PT_MIN_BLOCKS
#define PT_MIN_BLOCKS 3
#endif
__global__ void __launch_bounds__(PT_THREADS, PT_MIN_BLOCKS) k_pretok_count(PretokParams P) {
    __shared__ TileSmem S;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    u64* ck0 = (u64*)dyn_smem;
    u64* ck1 = ck0 + PT_CACHE_N;
    uint32_t* cc = (uint32_t*)(ck1 + PT_CACHE_N);
    for (int i = threadIdx.x; i < PT_CACHE_N; i += PT_THREADS) { ck0[i] = 0; ck1[i] = 0; cc[i] = 0; }
    init_tile_smem(S);
    const int tid = threadIdx.x;
    u64 my_tok = 0, my_us = 0, my_ul = 0, my_ub = 0;
    uint32_t phase[2] = {0, 0};
    // list mode: the items are (absolute tile, own_lo, own_hi) triples written by k_pretok_warp
    i64 n_items = P.n_tiles;
    if (P.list_mode) {
        n_items = P.stats[ST_SLOW_N];
        if (n_items > P.work_cap) { if (tid == 0 && blockIdx.x == 0) P.stats[ST_TABLE_FULL] = 4; n_items = P.work_cap; }
    }
    i64 item = blockIdx.x;
    if (item < n_items && tid == 0) tile_issue_load(P, S, P.list_mode ? P.work[3 * item] - P.tile_base : item, 0);
    int buf = 0;
    for (; item < n_items; item += gridDim.x, buf ^= 1) {
        const i64 next = item + gridDim.x;
        if (next < n_items && tid == 0) tile_issue_load(P, S, P.list_mode ? P.work[3 * next] - P.tile_base : next, buf ^ 1);
        const i64 tile = P.list_mode ? P.work[3 * item] - P.tile_base : item;
        const i64 own_lo = P.list_mode ? P.work[3 * item + 1] : P.own_lo, own_hi = P.list_mode ? P.work[3 * item + 2] : P.own_hi;
        mbar_wait(&S.bar[buf], phase[buf]); phase[buf] ^= 1;
        bool has_cut;
        tile_scan(P, S, tile, buf, &has_cut);

        const uint8_t* txt = S.txt[buf];
        const i64 g0 = (P.tile_base + tile) * PT_TILE - PT_HL;
        const int ntok = S.ntok_own, ntot = S.ntok_total;
        // cache misses are collected and sent to the global table together, so that their
        // probe latencies overlap instead of adding up
        u64 mk0[PT_MAXMISS], mk1[PT_MAXMISS], mh[PT_MAXMISS];
        int nmiss = 0;
        for (int k = tid; k < ntok; k += PT_THREADS) {
            int s = S.tokpos[k];
            i64 gpos = g0 + s;
            if (gpos < own_lo || gpos >= own_hi) continue;
            if (P.mode == 1 && P.n_sp > 0 && ((S.recw[(s >> 5) + 4] >> (s & 31)) & 1)) continue;   // encode: specials are not words
            my_tok++;
            if (k + 1 >= ntot) {                                      // end not inside the window
                u64 idx = atomicAdd((u64*)&P.stats[ST_OVF_N], 1ULL);
                if ((i64)idx < P.ovf_cap) P.ovf_pos[idx] = gpos;
                continue;
            }
            int len = (int)S.tokpos[k + 1] - s;
            int created;
            if (len <= PT_SHORT_MAX) {
                u64 k0, k1;
                pack_short_key(txt, s, len, &k0, &k1);
                u64 h = short_hash(k0, k1);
                if (!cache_add(ck0, ck1, cc, h, k0, k1)) {
                    if (nmiss < PT_MAXMISS) {
#pragma unroll
                        for (int i = 0; i < PT_MAXMISS; i++) if (i == nmiss) { mk0[i] = k0; mk1[i] = k1; mh[i] = h; }
                        nmiss++;
                    } else {
                        if (short_insert_h(P.st, h, k0, k1, 1, &created) < 0) P.stats[ST_TABLE_FULL] = 1;
                        if (created) { my_us++; my_ub += len; }
                    }
                }
            } else {
                u64 h = 0;
                for (int j = 0; j < len; j++) h += long_hash_term(txt[s + j], j);
                if (long_insert(P.lent, P.lcap, P.text, long_hash_fix(h), gpos, len, 1, &created) < 0) P.stats[ST_TABLE_FULL] = 1;
                if (created) { my_ul++; my_ub += len; }
            }
        }
        {
            ulonglong2 kv[PT_MAXMISS];
            const u64 smask_ = (u64)P.scap - 1;
#pragma unroll
            for (int i = 0; i < PT_MAXMISS; i++) if (i < nmiss) kv[i] = __ldcg((const ulonglong2*)P.st.key(mh[i] & smask_));
#pragma unroll
            for (int i = 0; i < PT_MAXMISS; i++) {
                if (i >= nmiss) continue;
                if (kv[i].x == mk0[i] && kv[i].y == mk1[i]) atomicAdd((u64*)P.st.cnt(mh[i] & smask_), 1ULL);
                else {
                    int created;
                    if (short_insert_h(P.st, mh[i], mk0[i], mk1[i], 1, &created) < 0) P.stats[ST_TABLE_FULL] = 1;
                    if (created) { my_us++; my_ub += (u64)(mk0[i] >> 56); }
                }
            }
        }
        __syncthreads();     // everyone is done with txt[buf] and S before the next iteration reuses them
    }
    // flush the pre-aggregation cache
    __syncthreads();
    for (int i = tid; i < PT_CACHE_N; i += PT_THREADS) {
        u64 k1 = ck1[i];
        if (k1 == 0) continue;
        u64 k0 = ck0[i];
        int created;
        if (short_insert(P.st, k0, k1, (i64)cc[i], &created) < 0) P.stats[ST_TABLE_FULL] = 1;
        if (created) { my_us++; my_ub += (u64)(k0 >> 56); }
    }
    // block-level reduction of the statistics
    for (int o = 16; o > 0; o >>= 1) {
        my_tok += __shfl_xor_sync(0xffffffffu, my_tok, o); my_us += __shfl_xor_sync(0xffffffffu, my_us, o);
        my_ul += __shfl_xor_sync(0xffffffffu, my_ul, o); my_ub += __shfl_xor_sync(0xffffffffu, my_ub, o);
    }
    if ((tid & 31) == 0) {
        if (my_tok) atomicAdd((u64*)&P.stats[ST_NTOK], my_tok);
        if (my_us) atomicAdd((u64*)&P.stats[ST_UNIQ_SHORT], my_us);
        if (my_ul) atomicAdd((u64*)&P.stats[ST_UNIQ_LONG], my_ul);
        if (my_ub) atomicAdd((u64*)&P.stats[ST_UNIQ_BYTES], my_ub);
    }
}

// ---------------------------------------------------------------------------------
// very long pre-tokens (end beyond the tile window): one CTA per token
// ---------------------------------------------------------------------------------
// find the end of the token starting at s: the first token start > s (or the logical end)
__device__ i64 block_find_token_end(const PretokParams& P, i64 s, i64* sh_min) {
    GlobalText G{P.text, P.n, P.cuts, P.n_cuts, P.n_sp > 0 ? P.rec : nullptr, -1, P.mode};
    i64 base = s + 1;
    for (;;) {
        if (threadIdx.x == 0) *sh_min = INT64_MAX;
        __syncthreads();
        i64 p = base + threadIdx.x;
        bool st = p >= P.n ? true : is_token_start(G, p);
        if (st) atomicMin((i64*)sh_min, p < P.n ? p : P.n);
        __syncthreads();
        i64 m = *sh_min;
        __syncthreads();
        if (m != INT64_MAX) return m;
        base += blockDim.x;
    }
}

__device__ u64 block_long_hash(const uint8_t* text, i64 s, i64 len, u64* sh_acc) {
    u64 h = 0;
    for (i64 j = threadIdx.x; j < len; j += blockDim.x) h += long_hash_term(text[s + j], j);
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if (threadIdx.x == 0) *sh_acc = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) atomicAdd(sh_acc, h);
    __syncthreads();
    u64 r = *sh_acc;
    __syncthreads();
    return long_hash_fix(r);
}

// block-cooperative version of long_insert / long_find (find_only: no insertion, returns slot or -1)
__device__ i64 block_long_upsert(LongEntry* ent, i64 cap, const uint8_t* text, u64 h, i64 pos, i64 len, i64 add,
                                 bool find_only, int* created, i64* sh) {
    // sh[0] = state (0 advance, 1 retry, 2 compare, 3 done, 4 fail), sh[1] = slot, sh[2] = epos
    u64 mask = (u64)cap - 1;
    if (threadIdx.x == 0) { sh[1] = (i64)(h & mask); sh[3] = 0; }
    *created = 0;
    for (int probes = 0; probes < 1 << 20; probes++) {
        __syncthreads();
        if (threadIdx.x == 0) {
            u64 slot = (u64)sh[1];
            u64* hp = &ent[slot].h;
            u64 cur = *(volatile u64*)hp;
            int state = 0;
            if (cur == 0) {
                if (find_only) state = 4;
                else {
                    cur = atomicCAS(hp, 0ULL, 1ULL);
                    if (cur == 0) {
                        ent[slot].pos = pos; ent[slot].len = len; ent[slot].count = 0;
                        __threadfence();
                        atomicExch(hp, h);
                        atomicAdd((u64*)&ent[slot].count, (u64)add);
                        sh[3] = 1; state = 3;
                    }
                }
            }
            if (state == 0) {
                if (cur == 1) state = 1;
                else if (cur == h) {
                    __threadfence();
                    i64 elen = *(volatile i64*)&ent[slot].len, epos = *(volatile i64*)&ent[slot].pos;
                    if (elen == len) { sh[2] = epos; state = epos == pos ? 5 : 2; }
                }
            }
            sh[0] = state;
        }
        __syncthreads();
        int state = (int)sh[0];
        i64 slot = sh[1];
        if (state == 3) { *created = (int)sh[3]; return slot; }
        if (state == 4) return -1;
        if (state == 1) continue;
        bool match = state == 5;
        if (state == 2) {
            i64 epos = sh[2];
            int diff = 0;
            for (i64 j = threadIdx.x; j < len && !diff; j += blockDim.x) if (text[epos + j] != text[pos + j]) diff = 1;
            match = !__syncthreads_or(diff);
        }
        if (match) {
            if (threadIdx.x == 0 && !find_only) atomicAdd((u64*)&ent[slot].count, (u64)add);
            return slot;
        }
        __syncthreads();
        if (threadIdx.x == 0) sh[1] = (i64)(((u64)slot + 1) & mask);
    }
    return -1;
}

__global__ void __launch_bounds__(256) k_long_tokens(PretokParams P, i64 n_ovf) {
    __shared__ i64 sh_min; __shared__ u64 sh_acc; __shared__ i64 sh[4];
    for (i64 t = blockIdx.x; t < n_ovf; t += gridDim.x) {
        i64 s = P.ovf_pos[t];
        i64 e = block_find_token_end(P, s, &sh_min);
        i64 len = e - s;
        u64 h = block_long_hash(P.text, s, len, &sh_acc);
        int created;
        i64 slot = block_long_upsert(P.lent, P.lcap, P.text, h, s, len, 1, false, &created, sh);
        if (threadIdx.x == 0) {
            if (slot < 0) P.stats[ST_TABLE_FULL] = 1;
            if (created) { atomicAdd((u64*)&P.stats[ST_UNIQ_LONG], 1ULL); atomicAdd((u64*)&P.stats[ST_UNIQ_BYTES], (u64)len); }
        }
        __syncthreads();
    }
}
