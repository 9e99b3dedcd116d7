// pretok_fast.cuh -- warp-autonomous pre-tokenise + count kernel (trainer mode, interior text).
//
// Same contract as k_pretok_count (pretok.cuh): replaces /root/reference/src/yet_another_bpe/
// trainer.py:146-170 (regex.findall of the GPT-2 pattern) + trainer.py:221-225 (word_freq).
//
// One persistent CTA per SM, PW_WARPS warps, and NO block-level synchronisation in the main loop:
// every warp owns a private double-buffered TMA pipeline over 992-byte chunks of the corpus
//   lane l  scans the 32-byte segment l of the chunk in registers (segment_scan, pretok.cuh);
//           lane 31's segment is the look-ahead that closes the chunk's last pre-token
//   warp    prefix-sums the start counts, writes the start offsets to its own shared-memory list
//   lane k  packs pre-token k into the 16-byte key, hashes it and counts it in the CTA-wide
//           shared-memory cache (one LDS.128 + compare + shared atomic on a hit)
//   misses  (cache full / long pre-tokens) go to a per-warp queue that is drained to the global
//           tables four independent probes per lane at a time
// Chunks that touch a hard cut, the ends of the text or the edge of the owned range are not handled
// here: they are appended to P.work and counted by k_pretok_count in list mode (generic path).
#pragma once

#include "pretok.cuh"

#define PW_CH 992                          // bytes owned per chunk (31 segments)
#define PW_HL 32                           // left context in the window (= PT_HL: segment_scan uses it)
#define PW_WIN 1088                        // 32 + 992 + 32 (look-ahead segment) + 32 (slack for its own look-ahead)
#ifndef PW_WARPS
#define PW_WARPS 24
#endif
#define PW_THREADS (PW_WARPS * 32)
#ifndef PW_NC
#define PW_NC 4096                         // cache entries per CTA (power of two)
#endif
#define PW_NC_LOG2 (PW_NC == 4096 ? 12 : (PW_NC == 2048 ? 11 : (PW_NC == 8192 ? 13 : 10)))
#define PW_QCAP 96                         // per-warp miss queue
#define PW_TOKCAP 1024

static_assert(PW_HL == PT_HL, "segment_scan assumes PT_HL bytes of left context");
static_assert((PW_NC & (PW_NC - 1)) == 0, "cache size must be a power of two");

struct WarpSmem {
    alignas(128) uint8_t txt[2][PW_WIN];
    alignas(16) ulonglong2 queue[PW_QCAP];
    uint16_t tokpos[PW_TOKCAP];
    alignas(8) uint64_t bar[2];
    uint8_t pad_[96];
};
static_assert(sizeof(WarpSmem) % 128 == 0, "warp areas stay 128-byte aligned");

#define PW_SMEM_BYTES (PW_NC * 20 + 256 + PW_WARPS * (int)sizeof(WarpSmem))

// no cut inside [g0, g1]?  (all lanes ask the same question: the loads broadcast)
__device__ __forceinline__ bool pw_no_cut(const PretokParams& P, i64 g0, i64 g1) {
    if (P.n_cuts == 0) return true;
    int lo = 0, hi = P.n_cuts;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (P.cuts[mid] < g0) lo = mid + 1; else hi = mid; }
    return !(lo < P.n_cuts && P.cuts[lo] <= g1);
}

__device__ __forceinline__ bool pw_chunk_is_fast(const PretokParams& P, i64 c) {
    const i64 A = c * PW_CH, g0 = A - PW_HL;
    if (A < P.own_lo || A + PW_CH > P.own_hi) return false;
    if (g0 <= 0 || g0 + PW_WIN >= P.n) return false;
    return pw_no_cut(P, g0, g0 + PW_WIN);
}

// hand the owned part of chunk c to the generic kernel, split at its 8 KiB tile boundaries
__device__ void pw_emit_slow(const PretokParams& P, i64 c) {
    i64 lo = c * PW_CH, hi = lo + PW_CH;
    if (lo < P.own_lo) lo = P.own_lo;
    if (hi > P.own_hi) hi = P.own_hi;
    if (lo >= hi) return;
    for (i64 t = lo / PT_TILE; t <= (hi - 1) / PT_TILE; t++) {
        const i64 a = lo > t * PT_TILE ? lo : t * PT_TILE, b = hi < (t + 1) * PT_TILE ? hi : (t + 1) * PT_TILE;
        const i64 idx = (i64)atomicAdd((u64*)&P.stats[ST_SLOW_N], 1ULL);
        if (idx < P.work_cap) { P.work[3 * idx] = t; P.work[3 * idx + 1] = a; P.work[3 * idx + 2] = b; }
    }
}

// drain the warp's miss queue into the global tables; the key loads of four entries per lane are in
// flight together so that their latencies overlap
__device__ void pw_drain(const PretokParams& P, const ulonglong2* queue, int qn, u64& my_us, u64& my_ul, u64& my_ub) {
    const int lane = threadIdx.x & 31;
    const u64 smask = (u64)P.scap - 1;
    for (int base = 0; base < qn; base += 128) {
        ulonglong2 e[4], kv[4];
        u64 sl[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = base + u * 32 + lane;
            e[u].x = 0; e[u].y = 0; sl[u] = 0; kv[u].x = 0; kv[u].y = 0;
            if (i < qn) {
                e[u] = queue[i];
                if ((e[u].y >> 56) == 1) { sl[u] = short_hash(e[u].x, e[u].y) & smask; kv[u] = __ldcg(&P.skeys[sl[u]]); }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (e[u].y == 0) continue;
            int created = 0;
            if ((e[u].y >> 56) == 1) {
                if (kv[u].x == e[u].x && kv[u].y == e[u].y) atomicAdd((u64*)&P.scounts[sl[u]], 1ULL);
                else {
                    if (short_insert_h(P.skeys, P.scounts, P.scap, sl[u], e[u].x, e[u].y, 1, &created) < 0) P.stats[ST_TABLE_FULL] = 1;
                    if (created) { my_us++; my_ub += e[u].x >> 56; }
                }
            } else {
                const i64 pos = (i64)e[u].x, len = (i64)(e[u].y & 0xffffffffULL);
                u64 h = 0;
                for (i64 j = 0; j < len; j++) h += long_hash_term(P.text[pos + j], j);
                if (long_insert(P.lent, P.lcap, P.text, long_hash_fix(h), pos, len, 1, &created) < 0) P.stats[ST_TABLE_FULL] = 1;
                if (created) { my_ul++; my_ub += (u64)len; }
            }
        }
    }
}

__global__ void __launch_bounds__(PW_THREADS, 1) k_pretok_warp(PretokParams P, i64 c_lo, i64 c_hi) {
    extern __shared__ __align__(128) unsigned char pw_smem[];
    uint4* ckeys = (uint4*)pw_smem;                                   // 16-byte keys (k0 = x,y  k1 = z,w)
    uint32_t* ccnt = (uint32_t*)(pw_smem + PW_NC * 16);               // bit 31 = claimed, low bits = occurrences
    uint4* lut = (uint4*)(pw_smem + PW_NC * 20);                      // byte masks per pre-token length
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpSmem& W = *(WarpSmem*)(pw_smem + PW_NC * 20 + 256 + warp * sizeof(WarpSmem));

    for (int i = threadIdx.x; i < PW_NC; i += PW_THREADS) { ckeys[i] = make_uint4(0, 0, 0, 0); ccnt[i] = 0; }
    if (threadIdx.x < 16) {
        uint32_t m[4];
        for (int j = 0; j < 4; j++) {
            const int r = (int)threadIdx.x - 4 * j;
            m[j] = r >= 4 ? 0xffffffffu : (r <= 0 ? 0u : ((1u << (8 * r)) - 1u));
        }
        lut[threadIdx.x] = make_uint4(m[0], m[1], m[2], m[3]);
    }
    if (lane == 0) { mbar_init(&W.bar[0], 1); mbar_init(&W.bar[1], 1); mbar_fence_init(); }
    __syncthreads();

    u64 my_tok = 0, my_us = 0, my_ul = 0, my_ub = 0, my_hit = 0;
    int qn = 0;
    uint32_t phase0 = 0, phase1 = 0;
    const i64 nw = (i64)gridDim.x * PW_WARPS;
    const uint32_t lt_mask = (1u << lane) - 1u;

    // the next fast chunk at or after c (slow chunks on the way are handed to the generic kernel)
    auto advance = [&](i64 c) -> i64 {
        while (c < c_hi && !pw_chunk_is_fast(P, c)) { if (lane == 0) pw_emit_slow(P, c); c += nw; }
        return c;
    };
    auto issue = [&](i64 c, int buf) {
        mbar_expect_tx(&W.bar[buf], PW_WIN);
        tma_load_1d(W.txt[buf], P.text + (c * PW_CH - PW_HL), PW_WIN, &W.bar[buf]);
    };

    i64 c = advance(c_lo + (i64)blockIdx.x * PW_WARPS + warp);
    if (c < c_hi && lane == 0) issue(c, 0);
    int buf = 0;
    while (c < c_hi) {
        const i64 cn = advance(c + nw);
        if (cn < c_hi && lane == 0) issue(cn, buf ^ 1);
        if (buf == 0) { mbar_wait(&W.bar[0], phase0); phase0 ^= 1; } else { mbar_wait(&W.bar[1], phase1); phase1 ^= 1; }
        const uint8_t* txt = W.txt[buf];
        const i64 g0 = c * PW_CH - PW_HL;

        // ---- token starts of this lane's segment
        uint32_t na_unused;
        const uint32_t m = segment_scan(P, txt, g0, lane, &na_unused);
        const bool own = lane < 31;
        const int cnt = own ? __popc(m) : 0;
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        const int total = __shfl_sync(0xffffffffu, inc, 31);
        if (own) {
            int off = inc - cnt;
            uint32_t mm = m;
            const int xb = PW_HL + 32 * lane;
            while (mm) { const int bit = __ffs(mm) - 1; mm &= mm - 1; W.tokpos[off++] = (uint16_t)(xb + bit); }
        } else {
            W.tokpos[total] = m ? (uint16_t)(PW_HL + PW_CH + __ffs(m) - 1) : (uint16_t)0xFFFF;
        }
        __syncwarp();

        // ---- one lane per pre-token
        const uint32_t* tw = (const uint32_t*)txt;
        for (int base = 0; base < total; base += 32) {
            const int k = base + lane;
            bool miss = false;
            ulonglong2 ent; ent.x = 0; ent.y = 0;
            if (k < total) {
                const int s = W.tokpos[k], e = W.tokpos[k + 1];
                my_tok++;
                if (e == 0xFFFF) {                                    // ends beyond the look-ahead segment
                    const u64 idx = atomicAdd((u64*)&P.stats[ST_OVF_N], 1ULL);
                    if ((i64)idx < P.ovf_cap) P.ovf_pos[idx] = g0 + s;
                } else {
                    const int len = e - s;
                    if (len <= PT_SHORT_MAX) {
                        const int wi = s >> 2, sh = (s & 3) * 8;
                        const uint32_t a0 = tw[wi], a1 = tw[wi + 1], a2 = tw[wi + 2], a3 = tw[wi + 3], a4 = tw[wi + 4];
                        const uint4 mk = lut[len];
                        const uint32_t b0 = __funnelshift_r(a0, a1, sh) & mk.x, b1 = __funnelshift_r(a1, a2, sh) & mk.y;
                        const uint32_t b2 = __funnelshift_r(a2, a3, sh) & mk.z, b3 = __funnelshift_r(a3, a4, sh) & mk.w;
                        // key words (same layout as pack_short_key): k0 = 7 bytes | len << 56, k1 = 7 bytes | 1 << 56
                        const uint32_t kx = b0, ky = (b1 & 0x00ffffffu) | ((uint32_t)len << 24);
                        const uint32_t kz = (b1 >> 24) | (b2 << 8), kw = (b2 >> 24) | (b3 << 8) | (1u << 24);
                        const uint32_t h = short_hash_w(kx, ky, kz, kw);
                        uint32_t ci = h >> (32 - PW_NC_LOG2);
                        bool done = false;
#pragma unroll
                        for (int probe = 0; probe < 2 && !done; probe++) {
                            const uint4 a = ckeys[ci];
                            if (a.x == kx && a.y == ky && a.z == kz && a.w == kw) { atomicAdd(&ccnt[ci], 1u); done = true; }
                            else if (a.w == 0 && atomicCAS(&ccnt[ci], 0u, 0x80000001u) == 0u) { ckeys[ci] = make_uint4(kx, ky, kz, kw); done = true; }
                            else ci ^= 1u;
                        }
                        if (done) my_hit++;
                        else { miss = true; ent.x = (u64)kx | ((u64)ky << 32); ent.y = (u64)kz | ((u64)kw << 32); }
                    } else {
                        miss = true; ent.x = (u64)(g0 + s); ent.y = (2ULL << 56) | (u64)len;
                    }
                }
            }
            const uint32_t mm = __ballot_sync(0xffffffffu, miss);
            if (mm) {
                if (miss) W.queue[qn + __popc(mm & lt_mask)] = ent;
                qn += __popc(mm);
                if (qn > PW_QCAP - 32) { __syncwarp(); pw_drain(P, W.queue, qn, my_us, my_ul, my_ub); qn = 0; __syncwarp(); }
            }
        }
        __syncwarp();        // every lane is done with txt[buf] and tokpos before they are reused
        c = cn; buf ^= 1;
    }
    __syncwarp();
    pw_drain(P, W.queue, qn, my_us, my_ul, my_ub);

    // ---- flush the cache
    __syncthreads();
    for (int i = threadIdx.x; i < PW_NC; i += PW_THREADS) {
        const uint32_t cn = ccnt[i] & 0x7fffffffu;
        if (cn == 0) continue;
        const uint4 a = ckeys[i];
        const u64 k0 = (u64)a.x | ((u64)a.y << 32), k1 = (u64)a.z | ((u64)a.w << 32);
        int created;
        if (short_insert(P.skeys, P.scounts, P.scap, k0, k1, (i64)cn, &created) < 0) P.stats[ST_TABLE_FULL] = 1;
        if (created) { my_us++; my_ub += k0 >> 56; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        my_tok += __shfl_xor_sync(0xffffffffu, my_tok, o); my_us += __shfl_xor_sync(0xffffffffu, my_us, o);
        my_ul += __shfl_xor_sync(0xffffffffu, my_ul, o); my_ub += __shfl_xor_sync(0xffffffffu, my_ub, o);
        my_hit += __shfl_xor_sync(0xffffffffu, my_hit, o);
    }
    if (lane == 0) {
        if (my_tok) atomicAdd((u64*)&P.stats[ST_NTOK], my_tok);
        if (my_us) atomicAdd((u64*)&P.stats[ST_UNIQ_SHORT], my_us);
        if (my_ul) atomicAdd((u64*)&P.stats[ST_UNIQ_LONG], my_ul);
        if (my_ub) atomicAdd((u64*)&P.stats[ST_UNIQ_BYTES], my_ub);
        if (my_hit) atomicAdd((u64*)&P.stats[ST_CACHE_HIT], my_hit);
    }
}
