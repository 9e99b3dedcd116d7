// encode.cuh -- batched encode (K6).
//
// Replaces /root/reference/src/yet_another_bpe/tokenizer.py:152-308 (encode +
// _encode_word_impl).  Pipeline (every stage on the device):
//   1. k_pretok_count (mode = encode)         unique pre-tokens of the batch -> hash tables
//   2. k_compact_short/long                    -> flat word table (symbols = byte symbols)
//   3. k_encode_words                          BPE by rank on every UNIQUE word, in place
//   4. k_encode_tiles<false>                   ids per text tile   -> exclusive scan
//   5. k_encode_tiles<true>                    ids written in text order (+ per-document offsets)
// The reference's 8192-entry LRU cache (tokenizer.py:48,83-86) is result-neutral; its GPU
// analogue is step 3 running once per unique word.
#pragma once

#include "merge.cuh"

struct EncodeModel {
    const u64* mkey; const u64* mval; i64 mcap;     // (sym_a, sym_b) -> rank << 32 | result symbol
    const int32_t* byte_sym;                        // 256 entries
    const int32_t* sym_out;                         // symbol -> vocab id (unk substituted, tokenizer.py:297-306)
    const int32_t* sp_ids;                          // special (priority order) -> vocab id or -1 (dropped)
    int consistent;                                 // merge list is creation-ordered: batch rule is exact
};

__device__ __forceinline__ bool merge_rank(const EncodeModel& E, int32_t a, int32_t b, uint32_t* rank, int32_t* res) {
    u64 key = PAIR_KEY(a, b);
    u64 mask = (u64)E.mcap - 1, slot = mix64(key) & mask;
    for (;;) {
        u64 k = E.mkey[slot];
        if (k == key) { u64 v = E.mval[slot]; *rank = (uint32_t)(v >> 32); *res = (int32_t)(v & 0xffffffffu); return true; }
        if (k == 0) return false;
        slot = (slot + 1) & mask;
    }
}

// exact restatement of the heap loop: merge the adjacent pair with the smallest (rank, position)
__device__ int encode_word_thread(const EncodeModel& E, int32_t* s, int n) {
    for (int j = 0; j < n; j++) s[j] = E.byte_sym[s[j]];
    while (n > 1) {
        uint32_t br = 0xffffffffu; int bp = -1; int32_t bres = 0;
        for (int j = 0; j + 1 < n; j++) {
            uint32_t r; int32_t res;
            if (merge_rank(E, s[j], s[j + 1], &r, &res) && r < br) { br = r; bp = j; bres = res; }
        }
        if (bp < 0) break;
        if (E.consistent) {
            // all left->right non-overlapping occurrences of the rank-br pair (equivalent)
            int32_t x = s[bp], y = s[bp + 1];
            int o = bp, j = bp;
            while (j < n) {
                if (j + 1 < n && s[j] == x && s[j + 1] == y) { s[o++] = bres; j += 2; }
                else s[o++] = s[j++];
            }
            n = o;
        } else {
            s[bp] = bres;
            for (int j = bp + 1; j + 1 < n; j++) s[j] = s[j + 1];
            n--;
        }
    }
    return n;
}

#define ENC_LONG_WORD 192

// block-cooperative version for long words; scratch = one int32 per symbol (sym_word region)
__device__ int encode_word_block(const EncodeModel& E, int32_t* s, int32_t* scratch, int n, int* sh_i, u64* sh_u) {
    for (int j = threadIdx.x; j < n; j += blockDim.x) s[j] = E.byte_sym[s[j]];
    __syncthreads();
    for (;;) {
        if (n <= 1) break;
        // min (rank, position)
        u64 best = ~0ULL; int32_t bres = 0;
        for (int j = threadIdx.x; j + 1 < n; j += blockDim.x) {
            uint32_t r; int32_t res;
            if (merge_rank(E, s[j], s[j + 1], &r, &res)) { u64 key = ((u64)r << 32) | (uint32_t)j; if (key < best) { best = key; bres = res; } }
        }
        if (threadIdx.x == 0) *sh_u = ~0ULL;
        __syncthreads();
        if (best != ~0ULL) atomicMin(sh_u, best);
        __syncthreads();
        u64 gbest = *sh_u;
        if (gbest == ~0ULL) break;
        if (best == gbest) sh_i[0] = bres;
        __syncthreads();
        bres = sh_i[0];
        int bp = (int)(gbest & 0xffffffffu);
        int32_t x = s[bp], y = s[bp + 1];
        __syncthreads();
        if (!E.consistent || x == y) {
            if (!E.consistent) {
                // single occurrence: shift the tail left by one through the scratch buffer
                for (int j = bp + 2 + threadIdx.x; j < n; j += blockDim.x) scratch[j] = s[j];
                __syncthreads();
                for (int j = bp + 2 + threadIdx.x; j < n; j += blockDim.x) s[j - 1] = scratch[j];
                if (threadIdx.x == 0) s[bp] = bres;
                n--;
            } else {
                // runs of the same symbol: sequential left->right pass
                if (threadIdx.x == 0) {
                    int o = bp, j = bp;
                    while (j < n) {
                        if (j + 1 < n && s[j] == x && s[j + 1] == y) { s[o++] = bres; j += 2; }
                        else s[o++] = s[j++];
                    }
                    sh_i[1] = o;
                }
                __syncthreads();
                n = sh_i[1];
            }
            __syncthreads();
            continue;
        }
        // x != y: occurrences cannot overlap; parallel compaction via scratch
        if (threadIdx.x == 0) sh_i[1] = 0;
        __syncthreads();
        int carry = 0;
        for (int base = 0; base < n; base += blockDim.x) {
            int j = base + threadIdx.x;
            bool is_first = j + 1 < n && s[j] == x && s[j + 1] == y;
            bool is_second = j > 0 && j < n && s[j - 1] == x && s[j] == y;
            int keep = (j < n && !is_second) ? 1 : 0;
            int total;
            // block exclusive scan of keep
            int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
            int inc = keep;
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            if (lane == 31) sh_i[2 + wid] = inc;
            __syncthreads();
            int wbase = 0; total = 0;
            for (int k = 0; k < (int)blockDim.x / 32; k++) { int t = sh_i[2 + k]; if (k < wid) wbase += t; total += t; }
            if (keep) scratch[carry + wbase + inc - 1] = is_first ? bres : s[j];
            carry += total;
            __syncthreads();
        }
        for (int j = threadIdx.x; j < carry; j += blockDim.x) s[j] = scratch[j];
        n = carry;
        __syncthreads();
    }
    __syncthreads();
    return n;
}

struct EncodeOut {
    const int32_t* wsym; const i64* woff; const int32_t* wlen;
    const int32_t* sword; const int32_t* lword;
    i64* tile_count;        // per tile: ids produced (pass 1), then exclusive base (after the scan)
    int32_t* out_ids; i64 out_cap;
    i64* doc_off;           // n_cuts + 2 entries: id offset at each document start
};

// ids of token k of the tile: word id (or -1 for a special), returns count
template <bool WRITE>
__global__ void __launch_bounds__(PT_THREADS) k_encode_tiles(PretokParams P, EncodeModel E, EncodeOut O) {
    __shared__ TileSmem S;
    __shared__ i64 sh_min; __shared__ u64 sh_acc; __shared__ i64 sh[4];
    __shared__ int sh_ovf_k; __shared__ i64 sh_ovf_w;      // sh_ovf_w: lookup record (first id slot << 24 | id count) or -1
    __shared__ i64 sh_base;
    init_tile_smem(S);
    const int tid = threadIdx.x;
    uint32_t phase[2] = {0, 0};
    i64 tile = blockIdx.x;
    if (tile < P.n_tiles && tid == 0) tile_issue_load(P, S, tile, 0);
    int buf = 0;
    for (; tile < P.n_tiles; tile += gridDim.x, buf ^= 1) {
        i64 next = tile + gridDim.x;
        if (next < P.n_tiles && tid == 0) tile_issue_load(P, S, next, buf ^ 1);
        mbar_wait(&S.bar[buf], phase[buf]); phase[buf] ^= 1;
        bool has_cut;
        tile_scan(P, S, tile, buf, &has_cut);
        const uint8_t* txt = S.txt[buf];
        const i64 g0 = (P.tile_base + tile) * PT_TILE - PT_HL;
        const int ntok = S.ntok_own, ntot = S.ntok_total;

        // the last owned token may end beyond the window: resolve it with the whole block
        if (tid == 0) { sh_ovf_k = -1; sh_ovf_w = -1; }
        __syncthreads();
        if (ntok > 0 && ntok >= ntot) {
            int k = ntok - 1;
            i64 gpos = g0 + S.tokpos[k];
            const int sk = S.tokpos[k];
            if (gpos >= P.own_lo && gpos < P.own_hi && !(P.n_sp > 0 && ((S.recw[(sk >> 5) + 4] >> (sk & 31)) & 1))) {
                i64 e = block_find_token_end(P, gpos, &sh_min);
                u64 h = block_long_hash(P.text, gpos, e - gpos, &sh_acc);
                int created;
                i64 slot = block_long_upsert(P.lent, P.lcap, P.text, h, gpos, e - gpos, 0, true, &created, sh);
                if (tid == 0) { sh_ovf_k = k; sh_ovf_w = slot >= 0 ? P.lent[slot].count : -1; }
            }
            __syncthreads();
        }
        if (WRITE && tid == 0) sh_base = O.tile_count[tile];
        __syncthreads();

        i64 running = 0;       // ids emitted by earlier rounds of this tile
        for (int kbase = 0; kbase < ntok; kbase += PT_THREADS) {
            int k = kbase + tid;
            int cnt = 0; i64 info = -1; int32_t spid = -1;       // info: lookup record written by k_encode_finalize
            int s = 0; i64 gpos = 0; bool live = false;
            if (k < ntok) {
                s = S.tokpos[k]; gpos = g0 + s;
                live = gpos >= P.own_lo && gpos < P.own_hi;
            }
            if (live) {
                if (P.n_sp > 0 && ((S.recw[(s >> 5) + 4] >> (s & 31)) & 1)) {   // recognised special: its id, or dropped (tokenizer.py:177-181)
                    int sp = special_match(P.text, gpos, logical_end_after(P, gpos));
                    spid = sp >= 0 ? E.sp_ids[sp] : -1;
                    cnt = spid >= 0 ? 1 : 0;
                } else if (k == sh_ovf_k) {
                    info = sh_ovf_w;
                    cnt = info >= 0 ? (int)(info & 0xffffff) : 0;
                } else {
                    int len = (int)S.tokpos[k + 1] - s;
                    if (len <= PT_SHORT_MAX) {
                        u64 k0, k1;
                        pack_short_key(txt, s, len, &k0, &k1);
                        info = short_find_info(P.st, k0, k1);      // key and record in one round trip
                    } else {
                        u64 h = 0;
                        for (int j = 0; j < len; j++) h += long_hash_term(txt[s + j], j);
                        i64 slot = long_find(P.lent, P.lcap, P.text, long_hash_fix(h), gpos, len);
                        info = slot >= 0 ? P.lent[slot].count : -1;
                    }
                    cnt = info >= 0 ? (int)(info & 0xffffff) : 0;
                    if (info < 0) P.stats[ST_TABLE_FULL] = 2;    // cannot happen: every token was inserted in step 1
                }
            }
            int total;
            int off = block_exclusive_scan(cnt, S.scan_tmp, &total);
            if (WRITE && live) {
                i64 dst = sh_base + running + off;
                if (has_cut && O.doc_off) {                     // first token of a document?
                    int lo = 0, hi = P.n_cuts;
                    while (lo < hi) { int mid = (lo + hi) >> 1; if (P.cuts[mid] < gpos) lo = mid + 1; else hi = mid; }
                    if (lo < P.n_cuts && P.cuts[lo] == gpos) O.doc_off[lo + 1] = dst;
                }
                if (spid >= 0) { if (dst < O.out_cap) O.out_ids[dst] = spid; }
                else if (info >= 0) {
                    const int32_t* src = O.wsym + (info >> 24);           // already vocabulary ids (k_encode_finalize)
                    for (int j = 0; j < cnt; j++) if (dst + j < O.out_cap) O.out_ids[dst + j] = src[j];
                }
            }
            running += total;
        }
        if (!WRITE && tid == 0) O.tile_count[tile] = running;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// One pass instead of two: per tile, count the ids (probe per token), obtain the tile's base offset from its
// predecessors (decoupled look-back over a status word per tile), write the ids.  The tile is scanned and staged once; the
// second probe of a token re-reads a sector this SM touched microseconds ago.  Tiles are handed out by a ticket counter, so
// the predecessor of every tile is held by a CTA that is already running (no residency assumption, no deadlock).
//   state[t]            0 = nothing yet, (1 << 62) | ids of tile t (aggregate), (2 << 62) | ids of tiles 0..t (inclusive prefix)
//   state[n_tiles]      total number of ids (written by the CTA of the last tile)
//   state[n_tiles + 1]  ticket counter (zeroed by the caller, like the rest)
// The caller sizes out_ids by a guess; ids beyond out_cap are dropped and the total tells it to repeat with the exact size.
// ---------------------------------------------------------------------------------------------------------------------
#define EF_AGG (1ULL << 62)
#define EF_PREFIX (2ULL << 62)
#define EF_VALUE(x) ((x) & ((1ULL << 62) - 1))

__global__ void __launch_bounds__(PT_THREADS, 5) k_encode_tiles_fused(PretokParams P, EncodeModel E, EncodeOut O) {
    __shared__ TileSmem S;
    __shared__ i64 sh_min; __shared__ u64 sh_acc; __shared__ i64 sh[4];
    __shared__ int sh_ovf_k; __shared__ i64 sh_ovf_w;
    __shared__ i64 sh_base, sh_tile[2];
    init_tile_smem(S);
    const int tid = threadIdx.x;
    uint32_t phase[2] = {0, 0};
    volatile u64* state = (volatile u64*)O.tile_count;
    const int buf = 0;
    for (;;) {
        // A tile is taken only when its predecessor in THIS CTA is finished: holding a ticket while still working on the
        // previous tile would make every later tile wait for it (the look-back needs the tiles in ticket order).  The tile's
        // load (8.5 KB) is therefore not overlapped with the previous tile's work: ~1 us of ~60 per tile.
        if (tid == 0) {
            sh_tile[0] = (i64)atomicAdd((u64*)&O.tile_count[P.n_tiles + 1], 1ULL);
            if (sh_tile[0] < P.n_tiles) tile_issue_load(P, S, sh_tile[0], 0);
        }
        __syncthreads();
        const i64 tile = sh_tile[0];
        if (tile >= P.n_tiles) break;
        mbar_wait(&S.bar[buf], phase[buf]); phase[buf] ^= 1;
        bool has_cut;
        tile_scan(P, S, tile, buf, &has_cut);
        const uint8_t* txt = S.txt[buf];
        const i64 g0 = (P.tile_base + tile) * PT_TILE - PT_HL;
        const int ntok = S.ntok_own, ntot = S.ntok_total;
        if (tid == 0) { sh_ovf_k = -1; sh_ovf_w = -1; }
        __syncthreads();
        if (ntok > 0 && ntok >= ntot) {                   // the last owned token may end beyond the window
            int k = ntok - 1;
            i64 gpos = g0 + S.tokpos[k];
            const int sk = S.tokpos[k];
            if (gpos >= P.own_lo && gpos < P.own_hi && !(P.n_sp > 0 && ((S.recw[(sk >> 5) + 4] >> (sk & 31)) & 1))) {
                i64 e = block_find_token_end(P, gpos, &sh_min);
                u64 h = block_long_hash(P.text, gpos, e - gpos, &sh_acc);
                int created;
                i64 slot = block_long_upsert(P.lent, P.lcap, P.text, h, gpos, e - gpos, 0, true, &created, sh);
                if (tid == 0) { sh_ovf_k = k; sh_ovf_w = slot >= 0 ? P.lent[slot].count : -1; }
            }
            __syncthreads();
        }
        i64 running = 0;
#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {            // 0: count, look back; 1: write
            running = 0;
            for (int kbase = 0; kbase < ntok; kbase += PT_THREADS) {
                int k = kbase + tid;
                int cnt = 0; i64 info = -1; int32_t spid = -1;
                int s = 0; i64 gpos = 0; bool live = false;
                if (k < ntok) { s = S.tokpos[k]; gpos = g0 + s; live = gpos >= P.own_lo && gpos < P.own_hi; }
                if (live) {
                    if (P.n_sp > 0 && ((S.recw[(s >> 5) + 4] >> (s & 31)) & 1)) {
                        int sp = special_match(P.text, gpos, logical_end_after(P, gpos));
                        spid = sp >= 0 ? E.sp_ids[sp] : -1;
                        cnt = spid >= 0 ? 1 : 0;
                    } else if (k == sh_ovf_k) {
                        info = sh_ovf_w;
                        cnt = info >= 0 ? (int)(info & 0xffffff) : 0;
                    } else {
                        int len = (int)S.tokpos[k + 1] - s;
                        if (len <= PT_SHORT_MAX) {
                            u64 k0, k1;
                            pack_short_key(txt, s, len, &k0, &k1);
                            info = short_find_info(P.st, k0, k1);
                        } else {
                            u64 h = 0;
                            for (int j = 0; j < len; j++) h += long_hash_term(txt[s + j], j);
                            i64 slot = long_find(P.lent, P.lcap, P.text, long_hash_fix(h), gpos, len);
                            info = slot >= 0 ? P.lent[slot].count : -1;
                        }
                        cnt = info >= 0 ? (int)(info & 0xffffff) : 0;
                        if (info < 0) P.stats[ST_TABLE_FULL] = 2;
                    }
                }
                int total;
                int off = block_exclusive_scan(cnt, S.scan_tmp, &total);
                if (pass == 1 && live) {
                    i64 dst = sh_base + running + off;
                    if (has_cut && O.doc_off) {
                        int lo = 0, hi = P.n_cuts;
                        while (lo < hi) { int mid = (lo + hi) >> 1; if (P.cuts[mid] < gpos) lo = mid + 1; else hi = mid; }
                        if (lo < P.n_cuts && P.cuts[lo] == gpos) O.doc_off[lo + 1] = dst;
                    }
                    if (spid >= 0) { if (dst < O.out_cap) O.out_ids[dst] = spid; }
                    else if (info >= 0) {
                        const int32_t* src = O.wsym + (info >> 24);
                        for (int j = 0; j < cnt; j++) if (dst + j < O.out_cap) O.out_ids[dst + j] = src[j];
                    }
                }
                running += total;
            }
            if (pass == 0) {
                if (tid < 32) {                           // warp 0 looks back 32 predecessors at a time
                    i64 base = 0;
                    if (tile > 0) {
                        if (tid == 0) { state[tile] = EF_AGG | (u64)running; __threadfence(); }
                        for (i64 j0 = tile - 1;;) {
                            const i64 j = j0 - tid;
                            const u64 v = j >= 0 ? state[j] : EF_PREFIX;            // before the first tile: an empty prefix
                            const unsigned pm = __ballot_sync(0xffffffffu, (v & EF_PREFIX) != 0);
                            const unsigned zm = __ballot_sync(0xffffffffu, v == 0);
                            const int p = pm ? __ffs(pm) - 1 : 32;                     // nearest predecessor with an inclusive prefix
                            const unsigned need = p >= 32 ? 0xffffffffu : ((2u << p) - 1u);
                            if (zm & need) continue;                                   // somebody up to there has not published yet
                            i64 part = (need >> tid) & 1u ? (i64)EF_VALUE(v) : 0;
                            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                            base += part;
                            if (p < 32) break;
                            j0 -= 32;
                        }
                    }
                    if (tid == 0) {
                        state[tile] = EF_PREFIX | (u64)(base + running);
                        if (tile == P.n_tiles - 1) state[P.n_tiles] = (u64)(base + running);
                        sh_base = base;
                    }
                }
                __syncthreads();
            }
        }
        __syncthreads();     // everyone is done with txt[buf], S and sh_tile before they are reused
    }
}

// After the unique words are encoded: (1) symbols -> vocabulary ids in place, (2) every table slot gets the lookup
// record of its word (first id slot << 24 | id count) where the occurrence count used to be (not needed any more), so
// that the two tile passes go from a token to its ids with ONE probe instead of slot -> word -> length / offset.
__global__ void __launch_bounds__(256) k_encode_finalize_ids(EncodeModel E, int32_t* wsym, const i64* woff, const int32_t* wlen, i64 n_words) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 w = (i64)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride) {
        int32_t* s = wsym + woff[w];
        const int n = wlen[w];
        for (int j = 0; j < n; j++) s[j] = E.sym_out[s[j]];
    }
}
__global__ void __launch_bounds__(256) k_encode_finalize_slots(ShortTab st, i64 scap, const int32_t* sword, LongEntry* lent, i64 lcap,
                                                               const int32_t* lword, const i64* woff, const int32_t* wlen) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < scap + lcap; i += stride) {
        const bool is_short = i < scap;
        const int32_t wid = is_short ? sword[i] : lword[i - scap];
        if (wid < 0) continue;
        const i64 info = (woff[wid] << 24) | (i64)wlen[wid];
        if (is_short) *st.cnt((u64)i) = info; else lent[i - scap].count = info;
    }
}

// single-block exclusive scan of per-tile counts; total written to data[n]
__global__ void __launch_bounds__(1024) k_scan_tiles(i64* data, i64 n) {
    __shared__ i64 tmp[32]; __shared__ i64 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (i64 base = 0; base < n; base += blockDim.x) {
        i64 i = base + threadIdx.x;
        i64 v = i < n ? data[i] : 0, inc = v;
        int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        for (int o = 1; o < 32; o <<= 1) { i64 t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) tmp[wid] = inc;
        __syncthreads();
        i64 wbase = 0, tot = 0;
        for (int k = 0; k < (int)blockDim.x / 32; k++) { i64 t = tmp[k]; if (k < wid) wbase += t; tot += t; }
        if (i < n) data[i] = carry + wbase + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) data[n] = carry;
}
