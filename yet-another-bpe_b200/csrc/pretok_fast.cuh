// pretok_fast.cuh -- warp-autonomous pre-tokenise + count kernel (trainer and encode mode, interior text).
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
#define PW_WARPS 28
#endif
#define PW_THREADS (PW_WARPS * 32)
#ifndef PW_WARPS_HOT
#define PW_WARPS_HOT 24                    // the hot-table variant keeps a second probe in flight per lane: more registers, fewer warps
#endif
#ifndef PW_NC
#define PW_NC 4096                         // cache entries per CTA (power of two)
#endif
#define PW_NC_LOG2 (PW_NC == 4096 ? 12 : (PW_NC == 2048 ? 11 : (PW_NC == 8192 ? 13 : (PW_NC == 1024 ? 10 : 9))))
#ifndef PW_TWO_WAY
#define PW_TWO_WAY 0      // a second cache way costs more instructions per round than its extra hits save
#endif
#define PW_LQCAP 64                        // per-warp queue of long (> 14 byte) pre-tokens
#define PW_TOKCAP 1024

static_assert(PW_HL == PT_HL, "segment_scan assumes PT_HL bytes of left context");
static_assert((PW_NC & (PW_NC - 1)) == 0, "cache size must be a power of two");

struct WarpSmem {
    alignas(128) uint8_t txt[2][PW_WIN];
    u64 lqueue[PW_LQCAP];                  // long pre-tokens: global position | length << 40
    uint16_t tokpos[PW_TOKCAP];
    alignas(8) uint64_t bar[2];
    uint32_t recw[2][PW_WIN / 32 + 8];     // recognised-special bits of the window (+4 words of history, +4 slack)
    uint8_t pad_[128 - (16 + 2 * 4 * (PW_WIN / 32 + 8)) % 128];
};
static_assert(sizeof(WarpSmem) % 128 == 0, "warp areas stay 128-byte aligned");

#define PW_SMEM_BYTES (PW_NC * 20 + 256 + PW_WARPS * (int)sizeof(WarpSmem))
#define PW_SMEM_BYTES_HOT (PW_NC * 20 + 256 + PW_WARPS_HOT * (int)sizeof(WarpSmem))
#define PW_HOT_SHIFT 5                     // hot-table slot = (hash >> PW_HOT_SHIFT) & (cap - 1)

// first cut >= g (or INT64_MAX)
__device__ __forceinline__ i64 pw_next_cut(const PretokParams& P, i64 g) {
    int lo = 0, hi = P.n_cuts;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (P.cuts[mid] < g) lo = mid + 1; else hi = mid; }
    return lo < P.n_cuts ? P.cuts[lo] : INT64_MAX;
}

// hand the owned part of the chunk at byte offset A to the generic kernel, split at its 8 KiB tile boundaries
__device__ void pw_emit_slow(const PretokParams& P, i64 A) {
    i64 lo = A, hi = lo + PW_CH;
    if (lo < P.own_lo) lo = P.own_lo;
    if (hi > P.own_hi) hi = P.own_hi;
    if (lo >= hi) return;
    for (i64 t = lo / PT_TILE; t <= (hi - 1) / PT_TILE; t++) {
        const i64 a = lo > t * PT_TILE ? lo : t * PT_TILE, b = hi < (t + 1) * PT_TILE ? hi : (t + 1) * PT_TILE;
        const i64 idx = (i64)atomicAdd((u64*)&P.stats[ST_SLOW_N], 1ULL);
        if (idx < P.work_cap) { P.work[3 * idx] = t; P.work[3 * idx + 1] = a; P.work[3 * idx + 2] = b; }
    }
}

// long pre-tokens (15 bytes .. one chunk): one lane each, bytes read back from global memory (L2-hot)
__device__ void pw_drain_long(const PretokParams& P, const u64* lqueue, int qn, u64& my_ul, u64& my_ub) {
    for (int i = threadIdx.x & 31; i < qn; i += 32) {
        const u64 e = lqueue[i];
        const i64 pos = (i64)(e & ((1ULL << 40) - 1)), len = (i64)(e >> 40);
        if (len == 0) continue;                  // placeholder of an over-long pre-token (handled by k_long_tokens_dyn)
        u64 h = 0;
        for (i64 j = 0; j < len; j++) h += long_hash_term(P.text[pos + j], j);
        int created;
        if (long_insert(P.lent, P.lcap, P.text, long_hash_fix(h), pos, len, 1, &created) < 0) P.stats[ST_TABLE_FULL] = 1;
        if (created) { my_ul++; my_ub += (u64)len; }
    }
}

// HOT = true (tables far larger than the L2: millions of unique pre-tokens).  Random 32-byte DRAM sectors cap the plain kernel
// (its throughput does not move between 20 and 28 resident warps).  A direct-mapped table of P.hot.cap slots, small enough to
// stay in the L2, is probed first: a pre-token whose key sits there is counted there (flushed into the big table by
// k_hot_flush afterwards); only the rest goes on to the DRAM-sized table, through a SECOND probe in flight per lane -- no lane
// waits for DRAM unless both probes of a lane continue in the same round (rare).
template <bool HOT>
__global__ void __launch_bounds__(HOT ? PW_WARPS_HOT * 32 : PW_THREADS, 1) k_pretok_warp(PretokParams P, i64 c_lo, i64 c_hi) {
    constexpr int NWARPS = HOT ? PW_WARPS_HOT : PW_WARPS;
    constexpr int NTHREADS = NWARPS * 32;
    extern __shared__ __align__(128) unsigned char pw_smem[];
    uint4* ckeys = (uint4*)pw_smem;                                   // 16-byte keys (k0 = x,y  k1 = z,w)
    uint32_t* ccnt = (uint32_t*)(pw_smem + PW_NC * 16);               // bit 31 = claimed, low bits = occurrences
    uint4* lut = (uint4*)(pw_smem + PW_NC * 20);                      // byte masks per pre-token length
    // lane id and the warp's shared-memory area are made opaque to the compiler: otherwise it re-derives them from
    // threadIdx (S2R + integer ops) all over the inner loops instead of spending two registers (6 % of all instructions)
    int lane;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    const int warp = threadIdx.x >> 5;
    WarpSmem* Wp = (WarpSmem*)(pw_smem + PW_NC * 20 + 256 + warp * sizeof(WarpSmem));
    asm volatile("" : "+l"(Wp));
    __builtin_assume(__isShared(Wp));          // ...but keep the address space: without this every access through W is a GENERIC
                                               // LD / ST (the token-list reads alone were 19 % of all stall samples)
    WarpSmem& W = *Wp;

    // The cache starts from the hot set of the sizing sample when there is one (k_hot_select: for every cache index the
    // most frequent sample key that maps to it) -- claimed with count 0; the remaining slots are claimed first come.
    for (int i = threadIdx.x; i < PW_NC; i += NTHREADS) {
        const uint4 hk = P.hot_keys ? P.hot_keys[i] : make_uint4(0, 0, 0, 0);
        ckeys[i] = hk; ccnt[i] = hk.w ? 0x80000000u : 0u;
    }
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

    u64 my_us = 0, my_ul = 0, my_ub = 0;
    uint32_t my_tok = 0, my_miss = 0, n_special = 0;   // warp-uniform
    int lqn = 0;
    uint32_t phase0 = 0, phase1 = 0;
    const i64 stepA = (i64)gridDim.x * NWARPS * PW_CH, endA = c_hi * PW_CH;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const i64 nrec = (P.n + 63) / 32 + 1;
    const u64 smask = (u64)P.scap - 1;
    // chunks whose offset A lies in [fastA_lo, fastA_hi] are fully owned and away from both ends of the text
    const i64 fastA_lo = P.own_lo > PW_HL + 1 ? P.own_lo : PW_HL + 1;
    const i64 fastA_hi = (P.own_hi - PW_CH < P.n - 1 - PW_WIN + PW_HL) ? P.own_hi - PW_CH : P.n - 1 - PW_WIN + PW_HL;
    i64 next_cut = P.n_cuts ? -1 : INT64_MAX;      // first cut >= the current window start (refreshed lazily)

    // a global-table probe in flight: issued when a pre-token misses the cache, consumed one round later
    bool pend = false;
    uint32_t pkx = 0, pky = 0, pkz = 0, pkw = 0, pslot = 0;
    ulonglong2 pkv; pkv.x = 0; pkv.y = 0;
    // HOT: the probe above goes to the hot table (pslot keeps the HASH); a second one, to the big table, may be in flight too
    bool pendB = false;
    uint32_t bkx = 0, bky = 0, bkz = 0, bkw = 0, bslot = 0;
    ulonglong2 bkv; bkv.x = 0; bkv.y = 0;
    const u64 hmask = HOT ? (u64)P.hot.cap - 1 : 0;
    auto main_insert = [&](uint32_t slot, u64 k0, u64 k1, ulonglong2 seen, uint32_t len_tag) {
        int created;
        if (short_insert_seen(P.st, slot, k0, k1, 1, &created, seen) < 0) P.stats[ST_TABLE_FULL] = 1;     // `seen`: the probe already read this slot
        if (created) { my_us++; my_ub += len_tag; }
    };
    auto consume = [&]() {
        if (HOT && pendB) {                  // big table: hit, empty slot (insert), or somebody else's key (next slot, next round)
            const u64 k0 = (u64)bkx | ((u64)bky << 32), k1 = (u64)bkz | ((u64)bkw << 32);
            if (bkv.x == k0 && bkv.y == k1) { atomicAdd((u64*)P.st.cnt(bslot), 1ULL); pendB = false; }
            else if ((bkv.x | bkv.y) == 0 || (bkv.x == 0) != (bkv.y == 0)) { main_insert(bslot, k0, k1, bkv, bky >> 24); pendB = false; }
            else { bslot = (uint32_t)(((u64)bslot + 1) & smask); bkv = probe_ld16(P.st.key(bslot)); }
        }
        if (pend) {
            const u64 k0 = (u64)pkx | ((u64)pky << 32), k1 = (u64)pkz | ((u64)pkw << 32);
            if (!HOT) {
                if (pkv.x == k0 && pkv.y == k1) atomicAdd((u64*)P.st.cnt(pslot), 1ULL);
                else main_insert(pslot, k0, k1, pkv, pky >> 24);
            } else {
                const u64 hs = ((u64)pslot >> PW_HOT_SHIFT) & hmask;
                bool in_hot = pkv.x == k0 && pkv.y == k1;
                if (!in_hot && (pkv.x | pkv.y) == 0) {                  // free hot slot: claim it (one 16-byte CAS, once per slot)
                    const ulonglong2 old = atom_cas128(P.hot.key(hs), 0ULL, 0ULL, k0, k1);
                    in_hot = (old.x == 0 && old.y == 0) || (old.x == k0 && old.y == k1);
                }
                if (in_hot) atomicAdd((u64*)P.hot.cnt(hs), 1ULL);
                else if (!pendB) {                                      // on to the big table, consumed next round
                    pendB = true; bkx = pkx; bky = pky; bkz = pkz; bkw = pkw;
                    bslot = (uint32_t)((u64)pslot & smask);
                    bkv = probe_ld16(P.st.key(bslot));
                } else {                                                // both probes of this lane continue: the rare synchronous case
                    ulonglong2 none; none.x = ~0ULL; none.y = ~0ULL;
                    main_insert((uint32_t)((u64)pslot & smask), k0, k1, none, pky >> 24);
                }
            }
            pend = false;
        }
    };

    // the next fast chunk at or after offset A (slow chunks on the way are handed to the generic kernel)
    auto advance = [&](i64 A) -> i64 {
        for (; A < endA; A += stepA) {
            if (A >= fastA_lo && A <= fastA_hi) {
                if (next_cut < A - PW_HL) next_cut = pw_next_cut(P, A - PW_HL);
                if (next_cut > A - PW_HL + PW_WIN) break;
            }
            if (lane == 0) pw_emit_slow(P, A);
        }
        return A;
    };

    i64 A = advance((c_lo + (i64)blockIdx.x * NWARPS + warp) * PW_CH);
    if (A < endA) {
        if (lane == 0) { mbar_expect_tx(&W.bar[0], PW_WIN); tma_load_1d(W.txt[0], P.text + (A - PW_HL), PW_WIN, &W.bar[0]); }
        if (P.n_sp > 0) {
            const i64 wi = ((A - PW_HL) >> 5) - 4 + lane;
            W.recw[0][lane] = wi >= 0 ? P.rec[wi] : 0u;
            if (lane < PW_WIN / 32 + 8 - 32) W.recw[0][32 + lane] = wi + 32 < nrec ? P.rec[wi + 32] : 0u;
        }
    }
    int buf = 0;
    while (A < endA) {
        const i64 An = advance(A + stepA);
        uint32_t nr0 = 0, nr1 = 0;                 // next window's recognised-special bits: loaded now, stored after this chunk
        if (An < endA) {
            if (lane == 0) { mbar_expect_tx(&W.bar[buf ^ 1], PW_WIN); tma_load_1d(W.txt[buf ^ 1], P.text + (An - PW_HL), PW_WIN, &W.bar[buf ^ 1]); }
            if (P.n_sp > 0) {
                const i64 wi = ((An - PW_HL) >> 5) - 4 + lane;
                nr0 = P.rec[wi];                   // wi >= 0: An > A >= PW_HL + 1 + stepA
                if (lane < PW_WIN / 32 + 8 - 32 && wi + 32 < nrec) nr1 = P.rec[wi + 32];
            }
        }
        if (buf == 0) { mbar_wait(&W.bar[0], phase0); phase0 ^= 1; } else { mbar_wait(&W.bar[1], phase1); phase1 ^= 1; }
        const uint8_t* txt = W.txt[buf];
        const i64 g0 = A - PW_HL;

        // ---- token starts of this lane's segment
        const uint32_t m = segment_scan(P, txt, W.recw[buf], g0, lane);
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
        my_tok += (uint32_t)total;

        // ---- one lane per pre-token
        const uint32_t* tw = (const uint32_t*)txt;
        for (int base = 0; base < total; base += 32) {
            const int k = base + lane;
            int s = 0, len = 0;
            if (k < total) { s = W.tokpos[k]; len = (int)W.tokpos[k + 1] - s; }      // sentinel 0xFFFF: len > 14
            if (P.mode == 1 && P.n_sp > 0) {             // encode: a recognised special is a separator, not a word
                const bool is_sp = k < total && ((W.recw[buf][(s >> 5) + 4] >> (s & 31)) & 1u);
                if (is_sp) len = 0;                      // neither short nor long below (length 0 is ignored by the long queue)
                n_special += __popc(__ballot_sync(0xffffffffu, is_sp));
            }
            uint32_t kx = 0, ky = 0, kz = 0, kw = 0, h = 0;
            bool miss = false;
            const bool is_short = (unsigned)(len - 1) < (unsigned)PT_SHORT_MAX;
            if (is_short) {
                const int wi = s >> 2, sh = (s & 3) * 8;
                const uint32_t a0 = tw[wi], a1 = tw[wi + 1], a2 = tw[wi + 2], a3 = tw[wi + 3], a4 = tw[wi + 4];
                const uint4 mk = lut[len];
                const uint32_t b0 = __funnelshift_r(a0, a1, sh) & mk.x, b1 = __funnelshift_r(a1, a2, sh) & mk.y;
                const uint32_t b2 = __funnelshift_r(a2, a3, sh) & mk.z, b3 = __funnelshift_r(a3, a4, sh) & mk.w;
                // key words (same layout as pack_short_key): k0 = 7 bytes | len << 56, k1 = 7 bytes | 1 << 56
                kx = b0; ky = (b1 & 0x00ffffffu) | ((uint32_t)len << 24);
                kz = __funnelshift_r(b1, b2, 24); kw = __funnelshift_r(b2, b3, 24) | (1u << 24);
                h = short_hash_w(kx, ky, kz, kw);
                const uint32_t ci = h >> (32 - PW_NC_LOG2);
                const uint4 a = ckeys[ci];
                if (((a.x ^ kx) | (a.y ^ ky) | (a.z ^ kz) | (a.w ^ kw)) == 0) atomicAdd(&ccnt[ci], 1u);
                else miss = true;
            }
            if (miss) {                            // second cache way, or claim an empty one
                uint32_t ci = h >> (32 - PW_NC_LOG2);
                if (ckeys[ci].w == 0 && atomicCAS(&ccnt[ci], 0u, 0x80000001u) == 0u) { ckeys[ci] = make_uint4(kx, ky, kz, kw); miss = false; }
#if PW_TWO_WAY
                else {
                    ci ^= 1u;
                    const uint4 a = ckeys[ci];
                    if (((a.x ^ kx) | (a.y ^ ky) | (a.z ^ kz) | (a.w ^ kw)) == 0) { atomicAdd(&ccnt[ci], 1u); miss = false; }
                    else if (a.w == 0 && atomicCAS(&ccnt[ci], 0u, 0x80000001u) == 0u) { ckeys[ci] = make_uint4(kx, ky, kz, kw); miss = false; }
                }
#endif
            }
            consume();                             // the probe issued one round ago has landed by now
            if (miss) {
                pend = true; pkx = kx; pky = ky; pkz = kz; pkw = kw;
                if (HOT) { pslot = h; pkv = probe_ld16(P.hot.key(((u64)h >> PW_HOT_SHIFT) & hmask)); }
                else { pslot = (uint32_t)(h & smask); pkv = probe_ld16(P.st.key(pslot)); }
            }
            my_miss += __popc(__ballot_sync(0xffffffffu, miss));
            // long (15 bytes .. one chunk) and over-long pre-tokens
            const bool is_long = k < total && !is_short && len > 0;
            const uint32_t lm = __ballot_sync(0xffffffffu, is_long);
            if (lm) {
                if (is_long) {
                    if (len + s == 0xFFFF) {                          // ends beyond the look-ahead segment
                        const u64 idx = atomicAdd((u64*)&P.stats[ST_OVF_N], 1ULL);
                        if ((i64)idx < P.ovf_cap) P.ovf_pos[idx] = g0 + s;
                        W.lqueue[lqn + __popc(lm & lt_mask)] = 0;     // placeholder (length 0: ignored by the drain)
                    } else W.lqueue[lqn + __popc(lm & lt_mask)] = (u64)(g0 + s) | ((u64)len << 40);
                }
                lqn += __popc(lm);
                my_miss += __popc(lm);
                if (lqn > PW_LQCAP - 32) { __syncwarp(); pw_drain_long(P, W.lqueue, lqn, my_ul, my_ub); lqn = 0; __syncwarp(); }
            }
        }
        if (P.n_sp > 0 && An < endA) {
            W.recw[buf ^ 1][lane] = nr0;
            if (lane < PW_WIN / 32 + 8 - 32) W.recw[buf ^ 1][32 + lane] = nr1;
        }
        __syncwarp();        // every lane is done with txt[buf], recw[buf] and tokpos before they are reused
        A = An; buf ^= 1;
    }
    consume();
    if (HOT) consume();                      // a probe moved to the big table by the call above
    while (HOT && pendB) consume();          // ... or still walking a collision chain
    __syncwarp();
    pw_drain_long(P, W.lqueue, lqn, my_ul, my_ub);

    // ---- flush the cache
    __syncthreads();
    for (int i = threadIdx.x; i < PW_NC; i += NTHREADS) {
        const uint32_t cn = ccnt[i] & 0x7fffffffu;
        if (cn == 0) continue;
        const uint4 a = ckeys[i];
        const u64 k0 = (u64)a.x | ((u64)a.y << 32), k1 = (u64)a.z | ((u64)a.w << 32);
        int created;
        if (short_insert(P.st, k0, k1, (i64)cn, &created) < 0) P.stats[ST_TABLE_FULL] = 1;
        if (created) { my_us++; my_ub += k0 >> 56; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        my_us += __shfl_xor_sync(0xffffffffu, my_us, o);
        my_ul += __shfl_xor_sync(0xffffffffu, my_ul, o); my_ub += __shfl_xor_sync(0xffffffffu, my_ub, o);
    }
    if (lane == 0) {
        if (my_tok > n_special) atomicAdd((u64*)&P.stats[ST_NTOK], (u64)(my_tok - n_special));
        if (my_us) atomicAdd((u64*)&P.stats[ST_UNIQ_SHORT], my_us);
        if (my_ul) atomicAdd((u64*)&P.stats[ST_UNIQ_LONG], my_ul);
        if (my_ub) atomicAdd((u64*)&P.stats[ST_UNIQ_BYTES], my_ub);
        if (my_tok > my_miss + n_special) atomicAdd((u64*)&P.stats[ST_CACHE_HIT], (u64)(my_tok - my_miss - n_special));
    }
}

// ---- hot table -> big table (after every k_pretok_warp<true>): counts move over, the claimed keys stay for the next call ----
__global__ void __launch_bounds__(256) k_hot_flush(PretokParams P) {
    u64 my_us = 0, my_ub = 0;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < P.hot.cap; i += stride) {
        const ulonglong2 kv = *(const ulonglong2*)P.hot.key((u64)i);
        if (kv.y == 0) continue;
        const i64 cnt = *P.hot.cnt((u64)i);
        if (cnt == 0) continue;
        *P.hot.cnt((u64)i) = 0;
        int created;
        if (short_insert(P.st, kv.x, kv.y, cnt, &created) < 0) P.stats[ST_TABLE_FULL] = 1;
        if (created) { my_us++; my_ub += kv.x >> 56; }
    }
    for (int o = 16; o > 0; o >>= 1) { my_us += __shfl_xor_sync(0xffffffffu, my_us, o); my_ub += __shfl_xor_sync(0xffffffffu, my_ub, o); }
    if ((threadIdx.x & 31) == 0) {
        if (my_us) atomicAdd((u64*)&P.stats[ST_UNIQ_SHORT], my_us);
        if (my_ub) atomicAdd((u64*)&P.stats[ST_UNIQ_BYTES], my_ub);
    }
}

// ---- hot set for the cache, from the table of the sizing sample ---------------------------------------------------
// pass 1: best[ci] = max over sample keys with cache index ci of (count << 24 | low slot bits); pass 2: the winner's key
__global__ void __launch_bounds__(256) k_hot_select(const ShortTab ST, u64* best, uint4* hot_keys, int pass) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < ST.cap; i += stride) {
        const ulonglong2 kv = *(const ulonglong2*)ST.key((u64)i);
        if (kv.y == 0) continue;
        const i64 cnt = *ST.cnt((u64)i);
        if (cnt < 2) continue;                                   // a word seen once in the sample is not worth a slot
        const uint32_t kx = (uint32_t)kv.x, ky = (uint32_t)(kv.x >> 32), kz = (uint32_t)kv.y, kw = (uint32_t)(kv.y >> 32);
        const uint32_t ci = short_hash_w(kx, ky, kz, kw) >> (32 - PW_NC_LOG2);
        const u64 tag = ((u64)cnt << 24) | ((u64)i & 0xffffffULL);
        if (pass == 0) atomicMax(&best[ci], tag);
        else if (best[ci] == tag) hot_keys[ci] = make_uint4(kx, ky, kz, kw);
    }
}
