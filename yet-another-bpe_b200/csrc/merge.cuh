// merge.cuh -- word table (K3), pair histogram (K4) and the persistent on-device merge loop (K5).
//
// Replaces /root/reference/src/yet_another_bpe/trainer.py:216-302 (_merge_loop).  Semantics
// (SURVEY.md 8(a) P5-P9, Appendix B): textbook BPE with
//   best = argmax (count, (left_bytes, right_bytes))              trainer.py:246
//   rewrite every affected word left->right, non-overlapping       trainer.py:276-285
//   always record the merge, new id only if the bytes are new     trainer.py:296-300
// Tokens are identified by their BYTES: two derivations of the same byte string share one id.
//
// One cooperative launch runs the whole loop; per merge: argmax over the active set
// (pairs with count >= T) -> rewrite of the candidate words (CSR postings + affected-word log)
// with incremental pair-count deltas.  See "merge loop" below for the two execution modes.
#pragma once

#include "pretok.cuh"

#ifndef ML_THREADS
#define ML_THREADS 1024
#endif
#ifndef ML_RW_TRACE
#define ML_RW_TRACE 0         // 1: thread 0 of CTA 0 times the stages of its own rewrite_batch chain (grid mode) into state[48..53]
#endif
#if ML_RW_TRACE
__device__ unsigned long long g_rw_clk[16];
#define SELT(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const long long _t = clock64(); atomicAdd(&g_rw_clk[8 + (k)], (unsigned long long)(_t - _s0)); _s0 = _t; } } while (0)
#define RWT(k, dep) do { if (!lc && blockIdx.x == 0 && threadIdx.x == 0) { long long _t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(_t) : "r"((int)(dep)) : "memory"); atomicAdd(&g_rw_clk[k], (unsigned long long)(_t - _t0)); _t0 = _t; } } while (0)
#else
#define RWT(k, dep)
#define SELT(k)
#endif
#ifndef ML_SEL_WHY
#define ML_SEL_WHY 0
#endif
#ifndef ML_TIMING
#define ML_TIMING 0           // 1: per-stage cycle counters of the leader loop in state[20..24] (costs registers)
#endif
#ifndef ML_TRACE
#define ML_TRACE 0            // > 0: thread 0 of CTA 0 writes a clock trace of leader merges [ML_TRACE, ML_TRACE + 8) to M.partial
#endif
#if ML_TRACE
__device__ long long* g_trace_row = nullptr;
#define ML_TRG(k) do { if (threadIdx.x == 0 && g_trace_row) g_trace_row[(k)] = clock64(); } while (0)
#define ML_TR(k) do { if (threadIdx.x == 0 && m >= ML_TRACE && m < ML_TRACE + 8) ((long long*)M.bsum)[512 + (m - ML_TRACE) * 24 + (k)] = clock64(); } while (0)
#else
#define ML_TR(k)
#define ML_TRG(k)
#endif
#if ML_TIMING
#define ML_CLOCK(v) long long v = clock64()
#define ML_T0(v) long long v = clock64()
#define ML_TACC(slot, from) do { if (threadIdx.x == 0) { long long _t = clock64(); sh_tacc[slot] += _t - (from); (from) = _t; } } while (0)
#else
#define ML_CLOCK(v)
#define ML_T0(v)
#define ML_TACC(slot, from)
#endif
#define PAIR_KEY(a, b) (0x8000000000000000ULL | ((u64)(uint32_t)(a) << 32) | (u64)(uint32_t)(b))
#define TOK_HASH_B 0x100000001b3ULL
// A posting / log entry carries the word AND its first symbol slot (fixed for the word's life): whoever picks a candidate
// can ask for the word's symbols in the same round trip as its header instead of after it.
#define POST_PACK(w, off) ((i64)(((u64)(uint32_t)(off) << 32) | (u64)(uint32_t)(w)))
#define POST_WORD(e) ((e) < 0 ? -1 : (int32_t)(uint32_t)((u64)(e) & 0xffffffffULL))
#define POST_OFF(e) ((i64)((u64)(e) >> 32))
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// merge-loop state slots (device int64[32])
#define MS_NMERGES 0
#define MS_NTOK 1
#define MS_ERROR 2
#define MS_ALOG_N 3
#define MS_ACT_N 4
#define MS_NPAIRS 6
#define MS_DONE 7
#define MS_POOL_USED 8
#define MS_REBUILDS 9
#define MS_TREBUILDS 10
#define MS_MAXCNT 11
#define MS_SCRATCH 12
#define MS_ACT_BASE 13
#define MS_LAST_REBUILD_M 14
#define MS_LEADER_MERGES 15
#define MS_GRID_MERGES 16
#define MS_LEADER_GEN 17
#define MS_TOP_N 18
#define MS_TOP_OVF 19
#define MS_T2 5
#define MS_LEADER_REASON 31
#define MS_TOP_REBUILDS 28
#define MS_BAR_COUNT 29
#define MS_BAR_GEN 30
#define MS_TIE_CNT 32        // count at which the top list overflowed (more than ML_TOP_N pairs tie for the maximum)
#define MS_TIE_LEFT 33       // merges left before the next attempt to rebuild the top list in that regime
#define MS_TOP_N_LIVE 34     // leader mode: current length of the top list (read by the prefetch helpers)
#define MS_LEADER_SMID 35
#define MS_T2_PA 36
// phase clocks of CTA 0 (cycles, always on: one thread reads %clock64 at phase boundaries)
#define MS_CLK_HIST 40       // initial pair histogram
#define MS_CLK_INDEX0 41     // first index build + active set
#define MS_CLK_TOPREB 42     // top-list rebuilds (incl. threshold steps)
#define MS_CLK_IDXREB 43     // later index rebuilds
#define MS_CLK_LEADER 44     // leader sessions (CTA 0 inside leader_loop)
#define MS_CLK_GRID 45       // grid-mode merges
#define MS_CLK_TOTAL 46
#define MS_N_TOPREB 47
#define MS_GRID_CLS 48        // grid merges by candidate-word count: [48..50] merges, [51..53] cycles; classes <= 2 368 (8-lane groups...), <= 18 944, more
#define ML_PHASE(slot, t0) do { if (gtid == 0) { const long long _t = clock64(); sh_phase[(slot) - 40] += _t - (t0); (t0) = _t; } } while (0)     // accumulated in shared memory, stored once at the end
#define MS_LEADER_ITERS 54    // leader iterations (an iteration merges a batch of 1 .. ML_BATCH_MAX pairs)
#define MS_LEADER_BATCHED 55  // merges done as members of a batch of two or more
#define MS_GRID_ITERS 56      // grid-mode iterations that merged a batch of two or more
#define MS_GRID_BATCHED 57    // merges done in those
#define MS_CLK_GB_SELECT 58   // grid-mode batches, cycles of CTA 0: selection | barrier 1 | ranges + lookups | own share of the rewrite | barrier 2
#define MS_CLK_GB_BAR1 59
#define MS_CLK_GB_RANGES 60
#define MS_CLK_GB_REWRITE 61
#define MS_CLK_GB_BAR2 62
#define MS_STATE_WORDS 64

#define ME_PAIR_TABLE_FULL 1
#define ME_TOK_POOL_FULL 2
#define ME_INTERNAL 4

// ---------------------------------------------------------------------------------
// K3: compact the pre-token tables into flat word arrays
// ---------------------------------------------------------------------------------
struct WordTable {
    int32_t* wsym; int32_t* sym_word; i64* woff; int32_t* wlen; i64* wcnt;
    int32_t* sword;      // short-table slot -> word id (or -1)
    int32_t* lword;      // long-table slot  -> word id (or -1)
    i64* counters;       // [0] n_words, [1] n_syms
};

// Word ids and symbol offsets are handed out with ONE pair of atomics per BLOCK ITERATION of 1 024 slots (per-thread counts of
// four slots -> block scan -> one atomicAdd pair by thread 0): the two global counters are hit cap / 1024 times instead of once
// per occupied slot (millions of unique pre-tokens) or per warp (a million same-address atomics made this scan of a 1 GB table
// run at 330 GB/s).
#define CS_PER_THREAD 4
__global__ void __launch_bounds__(256) k_compact_short(const ShortTab ST, i64 cap, WordTable W) {
    __shared__ int sh_w[8], sh_s[8];
    __shared__ i64 sh_base[2];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const i64 per_iter = 256 * CS_PER_THREAD, stride = (i64)gridDim.x * per_iter;
    for (i64 base = (i64)blockIdx.x * per_iter; base < cap; base += stride) {           // block-uniform trip count
        ulonglong2 kv[CS_PER_THREAD]; i64 oc[CS_PER_THREAD];
        int nw = 0, ns = 0;
#pragma unroll
        for (int j = 0; j < CS_PER_THREAD; j++) {
            const i64 i = base + threadIdx.x + 256 * j;
            kv[j].x = 0; kv[j].y = 0; oc[j] = 0;
            if (i < cap) { kv[j] = *(const ulonglong2*)ST.key((u64)i); oc[j] = *ST.cnt((u64)i); }
            if (kv[j].y != 0) { nw++; ns += (int)(kv[j].x >> 56); }
        }
        // block-exclusive scan of (words, symbols)
        int iw = nw, is = ns;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int tw = __shfl_up_sync(0xffffffffu, iw, o), ts = __shfl_up_sync(0xffffffffu, is, o);
            if (lane >= o) { iw += tw; is += ts; }
        }
        if (lane == 31) { sh_w[wid] = iw; sh_s[wid] = is; }
        __syncthreads();
        int bw = 0, bs = 0, tot_w = 0, tot_s = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { if (k < wid) { bw += sh_w[k]; bs += sh_s[k]; } tot_w += sh_w[k]; tot_s += sh_s[k]; }
        if (threadIdx.x == 0 && tot_w) {
            sh_base[0] = (i64)atomicAdd((u64*)&W.counters[0], (u64)tot_w);
            sh_base[1] = (i64)atomicAdd((u64*)&W.counters[1], (u64)tot_s);
        }
        __syncthreads();
        i64 wid0 = sh_base[0] + bw + iw - nw, off0 = sh_base[1] + bs + is - ns;
#pragma unroll
        for (int j = 0; j < CS_PER_THREAD; j++) {
            const i64 i = base + threadIdx.x + 256 * j;
            int32_t wd = -1;
            if (kv[j].y != 0) {
                const int len = (int)(kv[j].x >> 56);
                wd = (int32_t)wid0;
                W.woff[wd] = off0; W.wlen[wd] = len; W.wcnt[wd] = oc[j];
                for (int k = 0; k < len; k++) {
                    const int b = k < 7 ? (int)((kv[j].x >> (8 * k)) & 0xff) : (int)((kv[j].y >> (8 * (k - 7))) & 0xff);
                    W.wsym[off0 + k] = b; W.sym_word[off0 + k] = wd;
                }
                wid0++; off0 += len;
            }
            if (W.sword && i < cap) W.sword[i] = wd;
        }
        __syncthreads();                  // sh_base / sh_w are reused by the next iteration
    }
}

// long pre-tokens: every warp scans 32 entries at a time and copies the occupied ones with all its lanes
__global__ void __launch_bounds__(256) k_compact_long(const LongEntry* ent, i64 cap, const uint8_t* text, WordTable W) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    for (i64 base = (i64)blockIdx.x * blockDim.x; base < cap; base += stride) {
        const i64 i = base + threadIdx.x;
        LongEntry e; e.h = 0; e.pos = 0; e.len = 0; e.count = 0;
        if (i < cap) e = ent[i];
        const bool occ = e.h >= 2;
        uint32_t m = __ballot_sync(0xffffffffu, occ);
        int32_t wid = -1;
        while (m) {
            const int src = __ffs(m) - 1; m &= m - 1;
            const i64 len = __shfl_sync(0xffffffffu, e.len, src), pos = __shfl_sync(0xffffffffu, e.pos, src);
            const i64 cnt = __shfl_sync(0xffffffffu, e.count, src);
            i64 w0 = 0, off0 = 0;
            if (lane == 0) { w0 = (i64)atomicAdd((u64*)&W.counters[0], 1ULL); off0 = (i64)atomicAdd((u64*)&W.counters[1], (u64)len); }
            w0 = __shfl_sync(0xffffffffu, w0, 0); off0 = __shfl_sync(0xffffffffu, off0, 0);
            if (lane == 0) { W.woff[w0] = off0; W.wlen[w0] = (int32_t)len; W.wcnt[w0] = cnt; }
            for (i64 k = lane; k < len; k += 32) { W.wsym[off0 + k] = text[pos + k]; W.sym_word[off0 + k] = (int32_t)w0; }
            if (lane == src) wid = (int32_t)w0;
        }
        if (W.lword && i < cap) W.lword[i] = wid;
    }
}

// ---------------------------------------------------------------------------------
// merge loop
// ---------------------------------------------------------------------------------
// Index of "which words contain pair (p, q)":
//   * CSR postings built from the current words (rebuild_index), valid for every adjacency that
//     existed at rebuild time;
//   * every adjacency created later involves the token produced by the merge that created it,
//     so the words rewritten by merge m (the "affected log" segment of m) are a superset of the
//     words that contain any pair created by m.  Candidates for (p, q) = CSR postings of its slot
//     + the segments of the merges that produced p or q since the last rebuild.
// Duplicate candidates are filtered by a per-word stamp (grid mode) or a shared-memory claim set (leader
// mode); stale ones cost a look at the word and change nothing (the rewrite finds no site).
//
// Two execution modes, chosen uniformly by all CTAs after a grid barrier:
//   grid mode   every CTA takes part (argmax -> grid sync -> rewrite -> grid sync)
//   leader mode CTA 0 alone runs merges back to back with block barriers only, while the
//               other CTAs wait at one grid barrier; used while the active set and the
//               affected word lists are small (the common case after the first few hundred merges)
struct Best { i64 cnt; int32_t slot; int32_t a; int32_t b; int32_t pad; };
#ifndef ML_TOP_LOG2
#define ML_TOP_LOG2 9
#endif
#define ML_TOP_N (1 << ML_TOP_LOG2)          // entries of the top list: one per thread in the argmax (<= ML_THREADS)

#define ML_MAX_RANGES 12
#ifndef ML_LEADER_ITEMS_MAX
#define ML_LEADER_ITEMS_MAX 1024
#endif
#define ML_LEADER_BATCH 4096
#ifndef ML_LEADER_BATCH_ITEMS
#define ML_LEADER_BATCH_ITEMS 256      // a batch with more candidate words than this goes to the grid
#endif

struct MergeParams {
    // words
    int32_t* wsym; const int32_t* sym_word; i64 n_syms;
    int32_t* wslot;                              // per symbol slot: pair-table slot of (sym[j], sym[j+1])
    int32_t* newp;                               // n_syms entries: spill of the leader's per-merge list of new pairs
    const i64* woff; int32_t* wlen; const i64* wcnt; i64 n_words; int32_t* wstamp;
    // tokens (ids < n_base are prepared by the host: 256 bytes + specials)
    uint8_t* tok_bytes; i64 tok_bytes_cap; i64* tok_off; u64* tok_hash; u64* tok_pow;
    u64* tok_pre;                                // first 8 bytes of every token, big-endian, zero padded (tie-break fast path)
    u64* tset; i64 tset_cap; i64 max_tokens;
    // pairs
    u64* pkey; i64* pcnt; i64 pcap;
    uint32_t* ioff; uint32_t* icnt; i64* ipost; uint32_t* inact; uint32_t* intop; int32_t* act;      // ipost / alog_word entries: word | first symbol slot << 32
    int32_t* top_slot; u64* top_key; int* hist;
    // affected-word log: alog_word[seg_start[m] .. seg_end[m]) = words rewritten by merge m
    i64* alog_word; i64 alog_cap;
    int32_t* seg_start; int32_t* seg_end;      // per merge
    int32_t* merge_next;                        // per merge: previous merge (since rebuild) with the same product token
    int32_t* tok_first;                         // per token: latest merge since the last rebuild producing it, or -1
    int4* tok_head;                             // per token: {that merge, its segment start, end, the previous such merge}: one load instead of a chain
    Best* partial; i64* bsum;
    // outputs
    int32_t* merges; int32_t* merge_new; i64* state;
    i64 num_merges; i64 min_freq;
    i64 rebuild_every;                           // merges between index rebuilds (0: only when the affected-word log is full)
    i64 helper_min_syms;                         // prefetch helpers run when the word arrays have more symbol slots than this
    i64 helper_mode;                             // 0 off, 1 CTAs 1..ML_HELPERS, 2 the CTAs on the SMs next to the leader's (smid ^ 1, ^ 2)
    i64 batch_max;                               // merges per leader iteration: 0 = ML_BATCH_MAX, 1 = the reference's one-by-one loop
};

// Grid-wide barrier on two words of the state array (arrival counter + generation).  The kernel is
// launched cooperatively, so all CTAs are co-resident and spinning is safe.  Thread 0's fences order the
// CTA's earlier writes before the arrival and drop stale L1 lines after the release.
__device__ __forceinline__ void grid_barrier(const MergeParams& M) {
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile i64* gen = &M.state[MS_BAR_GEN];
        const i64 g = *gen;
        __threadfence();
        if (atomicAdd((u64*)&M.state[MS_BAR_COUNT], 1ULL) == (u64)gridDim.x - 1) {
            *(volatile i64*)&M.state[MS_BAR_COUNT] = 0;
            __threadfence();
            atomicAdd((u64*)&M.state[MS_BAR_GEN], 1ULL);
        } else {
            while (*gen == g) { }
        }
        __threadfence();
    }
    __syncthreads();
}

// expect_new: the key is very likely absent (it contains a token created by the current merge), so the probe
// starts with the CAS instead of a load: one L2 round trip less on the critical path of every merge.
// created_ctr: shared-memory counter of new pairs (leader mode), else the global MS_NPAIRS is bumped.
__device__ __forceinline__ i64 pair_upsert(const MergeParams& M, u64 key, bool expect_new = false, int* created_ctr = nullptr) {
    u64 mask = (u64)M.pcap - 1;
    u64 slot = mix64(key) & mask;
    for (i64 probes = 0; probes < M.pcap; probes++) {
        u64 k = expect_new ? 0ULL : M.pkey[slot];   // keys are write-once: a cached value is either right or empty
        if (k == 0) {
            k = atomicCAS(&M.pkey[slot], 0ULL, key);
            if (k == 0) {                           // load factor: checked once per merge
                if (created_ctr) atomicAdd(created_ctr, 1); else atomicAdd((u64*)&M.state[MS_NPAIRS], 1ULL);
                return (i64)slot;
            }
        }
        if (k == key) return (i64)slot;
        slot = (slot + 1) & mask;
    }
    atomicOr((u64*)&M.state[MS_ERROR], (u64)ME_PAIR_TABLE_FULL);
    return -1;
}

// Python bytes ordering: lexicographic unsigned, a proper prefix is smaller
__device__ int tok_cmp(const MergeParams& M, int32_t x, int32_t y) {
    if (x == y) return 0;
    i64 ox = M.tok_off[x], oy = M.tok_off[y];
    i64 lx = M.tok_off[x + 1] - ox, ly = M.tok_off[y + 1] - oy;
    i64 n = lx < ly ? lx : ly;
    for (i64 k = 0; k < n; k++) {
        int d = (int)M.tok_bytes[ox + k] - (int)M.tok_bytes[oy + k];
        if (d) return d;
    }
    return lx < ly ? -1 : (lx > ly ? 1 : 0);
}
// first 8 bytes, big-endian: for DIFFERENT prefixes the unsigned order of the prefixes is the bytes order of the tokens
// (a shorter token is zero padded: "ab" < "ab\x01"); EQUAL prefixes decide nothing (fall back to tok_cmp)
__device__ __forceinline__ u64 tok_prefix_of_bytes(const uint8_t* p, i64 len) {
    u64 v = 0;
    for (int k = 0; k < 8; k++) v = (v << 8) | (k < len ? (u64)p[k] : 0ULL);
    return v;
}
__device__ __forceinline__ u64 tok_prefix_concat(u64 pa, i64 la, u64 pb) { return la >= 8 ? pa : (pa | (pb >> (8 * la))); }
__device__ __forceinline__ int tok_cmp_pre(const MergeParams& M, int32_t x, u64 px, int32_t y, u64 py) {
    if (x == y) return 0;
    if (px != py) return px > py ? 1 : -1;
    return tok_cmp(M, x, y);
}
__device__ __forceinline__ bool best_gt(const MergeParams& M, const Best& p, const Best& q) {
    if (p.cnt != q.cnt) return p.cnt > q.cnt;
    if (p.slot == q.slot) return false;
    if (q.slot < 0) return true;
    if (p.slot < 0) return false;
    int r = tok_cmp(M, p.a, q.a);
    if (r) return r > 0;
    return tok_cmp(M, p.b, q.b) > 0;
}
__device__ __forceinline__ Best shfl_best(const Best& v, int o) {
    Best r;
    r.cnt = __shfl_xor_sync(0xffffffffu, v.cnt, o); r.slot = __shfl_xor_sync(0xffffffffu, v.slot, o);
    r.a = __shfl_xor_sync(0xffffffffu, v.a, o); r.b = __shfl_xor_sync(0xffffffffu, v.b, o); r.pad = 0;
    return r;
}
__device__ __forceinline__ Best shfl_best_from(const Best& v, int src) {
    Best r;
    r.cnt = __shfl_sync(0xffffffffu, v.cnt, src); r.slot = __shfl_sync(0xffffffffu, v.slot, src);
    r.a = __shfl_sync(0xffffffffu, v.a, src); r.b = __shfl_sync(0xffffffffu, v.b, src); r.pad = __shfl_sync(0xffffffffu, v.pad, src);
    return r;
}
__device__ Best block_best(const MergeParams& M, Best v, Best* sh) {
    for (int o = 16; o > 0; o >>= 1) { Best t = shfl_best(v, o); if (best_gt(M, t, v)) v = t; }
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    if (wid == 0) {
        Best t = lane < ML_THREADS / 32 ? sh[lane] : Best{0, -1, 0, 0, 0};
        for (int o = 16; o > 0; o >>= 1) { Best u = shfl_best(t, o); if (best_gt(M, u, t)) t = u; }
        if (lane == 0) sh[0] = t;
    }
    __syncthreads();
    Best r = sh[0];
    __syncthreads();
    return r;
}

// best over act[lo, hi) strided by this thread's position among `nthreads` cooperating threads
__device__ __forceinline__ Best scan_active(const MergeParams& M, i64 n, i64 first, i64 stride) {
    Best v{0, -1, 0, 0, 0};
    // 4 independent (act -> count, key) chains in flight per thread: the scan is latency bound.
    // __ldcg: counts are updated by atomics (L2); the leader loop has no fence that would drop a stale L1 line.
    for (i64 i0 = first; i0 < n; i0 += 4 * stride) {
        int32_t sl[4]; i64 cc[4]; u64 kk[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { i64 i = i0 + u * stride; sl[u] = i < n ? M.act[i] : -1; }
#pragma unroll
        for (int u = 0; u < 4; u++) { cc[u] = sl[u] >= 0 ? __ldcg(&M.pcnt[sl[u]]) : 0; kk[u] = sl[u] >= 0 ? __ldcg(&M.pkey[sl[u]]) : 0; }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (cc[u] <= 0 || cc[u] < v.cnt) continue;
            Best t{cc[u], sl[u], (int32_t)((kk[u] >> 32) & 0x7fffffff), (int32_t)(kk[u] & 0xffffffffu), 0};
            if (best_gt(M, t, v)) v = t;
        }
    }
    return v;
}

// block-wide best: max count first (cheap), the byte-wise tie-break only among equal counts
__device__ Best block_best_fast(const MergeParams& M, Best v, Best* sh, i64* sh_cnt) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    i64 c = v.slot >= 0 ? v.cnt : 0;
    for (int o = 16; o > 0; o >>= 1) { i64 t = __shfl_xor_sync(0xffffffffu, c, o); if (t > c) c = t; }
    if (lane == 0) sh_cnt[wid] = c;
    __syncthreads();
    i64 mx = lane < (int)(blockDim.x >> 5) ? sh_cnt[lane] : 0;
    for (int o = 16; o > 0; o >>= 1) { i64 t = __shfl_xor_sync(0xffffffffu, mx, o); if (t > mx) mx = t; }
    const bool cand = v.slot >= 0 && mx > 0 && v.cnt == mx;
    const int ncand = __syncthreads_count(cand);
    if (ncand == 0) return Best{0, -1, 0, 0, 0};
    if (ncand == 1) {
        if (cand) sh[0] = v;
        __syncthreads();
        Best r = sh[0];
        __syncthreads();
        return r;
    }
    return block_best(M, cand ? v : Best{0, -1, 0, 0, 0}, sh);
}

// grid-wide argmax over the active set; every block returns the same result
__device__ Best grid_argmax(const MergeParams& M, Best* sh) {
    Best v = scan_active(M, M.state[MS_ACT_N], (i64)blockIdx.x * blockDim.x + threadIdx.x, (i64)gridDim.x * blockDim.x);
    v = block_best(M, v, sh);
    if (threadIdx.x == 0) M.partial[blockIdx.x] = v;
    grid_barrier(M);
    Best w{0, -1, 0, 0, 0};
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) { Best t = M.partial[i]; if (best_gt(M, t, w)) w = t; }
    return block_best(M, w, sh);
}

__device__ void rebuild_active(const MergeParams& M, i64 T) {
    i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gstride = (i64)gridDim.x * blockDim.x;
    for (i64 i = gtid; i < (M.pcap + 31) / 32; i += gstride) M.inact[i] = 0;
    if (gtid == 0) { M.state[MS_ACT_N] = 0; M.state[MS_TREBUILDS]++; }
    grid_barrier(M);
    for (i64 s = gtid; s < M.pcap; s += gstride) {
        if (M.pkey[s] != 0 && M.pcnt[s] >= T) {
            atomicOr(&M.inact[s >> 5], 1u << (s & 31));
            i64 idx = (i64)atomicAdd((u64*)&M.state[MS_ACT_N], 1ULL);
            M.act[idx] = (int32_t)s;
        }
    }
    grid_barrier(M);

}

// CSR postings: for every pair slot the words that contain it (duplicates allowed)
__device__ void rebuild_index(const MergeParams& M, i64* sh_scan, i64 m_now) {
    i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gstride = (i64)gridDim.x * blockDim.x;
    for (i64 i = gtid; i < M.pcap; i += gstride) M.icnt[i] = 0;
    // forget the affected-log segments: the CSR built below covers everything
    for (i64 mm = M.state[MS_LAST_REBUILD_M] + gtid; mm < m_now; mm += gstride) { M.tok_first[M.merge_new[mm]] = -1; M.tok_head[M.merge_new[mm]].x = -1; }
    grid_barrier(M);
    for (i64 i = gtid; i + 1 < M.n_syms; i += gstride) {
        int32_t w = M.sym_word[i];
        i64 j = i - M.woff[w];
        if (j + 1 < M.wlen[w]) atomicAdd(&M.icnt[M.wslot[i]], 1u);
    }
    grid_barrier(M);
    // exclusive scan of icnt -> ioff, one contiguous chunk per block
    i64 chunk = (M.pcap + gridDim.x - 1) / gridDim.x;
    i64 lo = chunk * blockIdx.x, hi = lo + chunk < M.pcap ? lo + chunk : M.pcap;
    if (lo > hi) lo = hi;
    {
        i64 s = 0;
        for (i64 i = lo + threadIdx.x; i < hi; i += blockDim.x) s += M.icnt[i];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) sh_scan[0] = 0;
        __syncthreads();
        if ((threadIdx.x & 31) == 0) atomicAdd((u64*)&sh_scan[0], (u64)s);
        __syncthreads();
        if (threadIdx.x == 0) M.bsum[blockIdx.x] = sh_scan[0];
    }
    grid_barrier(M);
    {
        i64 base = 0;
        for (int b = 0; b < (int)blockIdx.x; b++) base += M.bsum[b];
        __syncthreads();
        for (i64 t0 = lo; t0 < hi; t0 += blockDim.x) {
            i64 i = t0 + threadIdx.x;
            i64 v = i < hi ? M.icnt[i] : 0;
            i64 inc = v;
            int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
            for (int o = 1; o < 32; o <<= 1) { i64 t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            if (lane == 31) sh_scan[1 + wid] = inc;
            __syncthreads();
            i64 wbase = 0, tot = 0;
            for (int k = 0; k < ML_THREADS / 32; k++) { i64 t = sh_scan[1 + k]; if (k < wid) wbase += t; tot += t; }
            if (i < hi) M.ioff[i] = (uint32_t)(base + wbase + inc - v);
            base += tot;
            __syncthreads();
        }
        if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) M.ioff[M.pcap] = (uint32_t)base;
    }
    grid_barrier(M);
    for (i64 i = gtid; i + 1 < M.n_syms; i += gstride) {
        int32_t w = M.sym_word[i];
        i64 j = i - M.woff[w];
        if (j + 1 < M.wlen[w]) { const int32_t s = M.wslot[i]; uint32_t r = atomicSub(&M.icnt[s], 1u) - 1; M.ipost[M.ioff[s] + r] = POST_PACK(w, i - j); }
    }
    if (gtid == 0) { M.state[MS_ALOG_N] = 0; M.state[MS_LAST_REBUILD_M] = m_now; M.state[MS_REBUILDS]++; }
    grid_barrier(M);
}

// Leader mode keeps the counts of the top-list pairs mirrored in shared memory (the leader is the only
// writer while it runs), so the per-merge argmax needs no global gather at all.
// Counts are kept as two 32-bit halves: shared memory has native 32-bit atomic adds, while a 64-bit add is a
// compare-and-swap loop (~145 cycles uncontended, far worse when every rewrite site hits the same entry).
struct LeaderMirror { uint32_t lo[ML_TOP_N], hi[ML_TOP_N]; uint8_t pending[ML_TOP_N]; int32_t mkey[2 * ML_TOP_N]; int16_t mval[2 * ML_TOP_N]; };
__device__ __forceinline__ i64 mirror_get(const LeaderMirror* lm, int idx) { return (i64)(((u64)lm->hi[idx] << 32) | lm->lo[idx]); }
__device__ __forceinline__ void mirror_set(LeaderMirror* lm, int idx, i64 v) { lm->lo[idx] = (uint32_t)(u64)v; lm->hi[idx] = (uint32_t)((u64)v >> 32); lm->pending[idx] = 0; }

__device__ __forceinline__ int mirror_find(const LeaderMirror* lm, int32_t slot) {
    uint32_t h = ((uint32_t)slot * 2654435761u) >> (31 - ML_TOP_LOG2);          // 2 * ML_TOP_N entries
    for (;;) {
        const int32_t k = lm->mkey[h];
        if (k == slot + 1) return lm->mval[h];
        if (k == 0) return -1;
        h = (h + 1) & (2 * ML_TOP_N - 1);
    }
}
__device__ __forceinline__ void mirror_insert(LeaderMirror* lm, int32_t slot, int idx) {
    uint32_t h = ((uint32_t)slot * 2654435761u) >> (31 - ML_TOP_LOG2);
    for (;;) {
        const int32_t old = atomicCAS(&lm->mkey[h], 0, slot + 1);
        if (old == 0) { lm->mval[h] = (int16_t)idx; return; }
        h = (h + 1) & (2 * ML_TOP_N - 1);
    }
}

// Leader-mode context (shared memory of CTA 0).  While the leader runs it is the only writer of the merge
// state, so the counters the rewrite bumps (active set, affected log, top list, new pairs) live here and
// are written back to M.state once, when the leader hands control back to the grid.
#define ML_DEDUPE_N (2 * ML_LEADER_ITEMS_MAX)
#define ML_NEWP_N 1024
struct LeaderCtx {
    LeaderMirror LM;
    int32_t tslot[ML_TOP_N]; u64 tkey[ML_TOP_N];
    u64 tpa[ML_TOP_N], tpb[ML_TOP_N];            // 8-byte prefixes of the two tokens of every entry (tie-break without global loads)
    u64 tha[ML_TOP_N], thb[ML_TOP_N], tpwb[ML_TOP_N];   // hash(a), hash(b), base^len(b): the merged token's hash without a round trip
    uint32_t tp0[ML_TOP_N], tplen[ML_TOP_N];     // CSR postings of the entry's slot (ioff is a DRAM-sized array: loaded with the mirror, off the critical path)
    int32_t dedupe[ML_DEDUPE_N];                 // word + 1: candidate words already taken by the current merge (0 = free)
    int32_t newp[ML_NEWP_N];                     // slots of the pairs created by the current merge (spill: M.newp)
    u64 newp_key[ML_NEWP_N];                     // and their keys (tokens known without waiting for the table)
    int top_n, top_ovf, act_n, alog_n, npairs_new, error, nnew;
    u64 t2pa;                                    // top-list threshold, second component (copy of state[MS_T2_PA])
};

__device__ __forceinline__ void mirror_add(LeaderCtx* lc, int32_t slot, i64 d) {
    if (!lc) return;
    const int idx = mirror_find(&lc->LM, slot);
    if (idx >= 0) {
        const uint32_t dlo = (uint32_t)(u64)d, dhi = (uint32_t)((u64)d >> 32);
        const uint32_t old = atomicAdd(&lc->LM.lo[idx], dlo);
        const uint32_t carry = (uint32_t)(((u64)old + dlo) >> 32);
        if (dhi + carry) atomicAdd(&lc->LM.hi[idx], dhi + carry);
    }
}

// true when this call took word w for the current merge (first claim wins); the set is cleared once per merge
__device__ __forceinline__ bool dedupe_claim(LeaderCtx* lc, int32_t w) {
    uint32_t h = (((uint32_t)w * 2654435761u) >> 8) & (ML_DEDUPE_N - 1);
    for (;;) {
        const int32_t old = atomicCAS(&lc->dedupe[h], 0, w + 1);
        if (old == 0) return true;
        if (old == w + 1) return false;
        h = (h + 1) & (ML_DEDUPE_N - 1);
    }
}

__device__ __forceinline__ void pair_sub(const MergeParams& M, int32_t slot, i64 f, LeaderCtx* lc) {
    atomicAdd((u64*)&M.pcnt[slot], (u64)(-f));     // the slot of every adjacency is cached in wslot: no probe
    mirror_add(lc, slot, -f);                      // (never the merged pair itself: its count is set to 0 once per merge)
}
// "Top list": every pair with count >= T2, kept in global memory (at most ML_TOP_N entries) and
// maintained by pair_add in both modes.  Scanning the whole active set every merge is bound by the
// gather rate of the SMs involved (one 128-byte line per cycle per SM); the top list keeps the
// per-merge scan to a few hundred entries.  It is rebuilt (grid-wide) when its best entry falls
// below T2 or when it overflows.  state[MS_T2]: > 0 valid threshold, 0 = rebuild needed,
// -1 = disabled for this merge (more than ML_TOP_N pairs tie for the maximum).
// When thousands of pairs tie at one count (small corpora, the tail of any run) a count threshold cannot cut the list to
// ML_TOP_N entries: the threshold then has a second component, the 8-byte prefix PA of the LEFT token -- the list holds every
// pair with count > T2, or count == T2 and tok_pre[left] >= PA.  The best pair orders by (count, left bytes, right bytes) and
// the prefix is monotone in the bytes, so this set is closed upwards in that order: its best entry is the global best as
// long as it is itself inside the set.  PA = 0 is the plain count threshold.
// Membership (active set: count >= T, top list: count >= T2) is updated by the add that CROSSES the
// threshold upwards (before < T <= after): exactly one add sees each crossing, so most adds end with the
// atomic that returns the new count and never touch the membership bitmaps.
__device__ __forceinline__ int32_t pair_add(const MergeParams& M, int32_t x, int32_t y, i64 f, i64 T, i64 T2, LeaderCtx* lc, bool expect_new) {
    const u64 key = PAIR_KEY(x, y);
    if (lc && expect_new) {
        // Leader mode, new product token: every pair that GAINS in this merge contains the new token, i.e. was
        // created by this merge.  The creator notes the slot; the count is added without waiting for the result
        // and the thresholds are checked once per new pair after the rewrite (leader_new_pairs).
        const u64 mask = (u64)M.pcap - 1;
        u64 slot = mix64(key) & mask;
        for (i64 probes = 0; probes < M.pcap; probes++) {
            const u64 k = atomicCAS(&M.pkey[slot], 0ULL, key);
            if (k == 0) {
                const int idx = atomicAdd(&lc->nnew, 1);
                if (idx < ML_NEWP_N) { lc->newp[idx] = (int32_t)slot; lc->newp_key[idx] = key; } else M.newp[idx - ML_NEWP_N] = (int32_t)slot;
                break;
            }
            if (k == key) break;
            slot = (slot + 1) & mask;
        }
        atomicAdd((u64*)&M.pcnt[slot], (u64)f);
        return (int32_t)slot;
    }
    i64 s = pair_upsert(M, key, expect_new, lc ? &lc->npairs_new : nullptr);
    if (s < 0) return 0;
    const i64 now = (i64)atomicAdd((u64*)&M.pcnt[s], (u64)f) + f;
    mirror_add(lc, (int32_t)s, f);
    if (now >= T && now - f < T) {
        const uint32_t bit = 1u << (s & 31);
        if (!(atomicOr(&M.inact[s >> 5], bit) & bit)) {
            const i64 idx = lc ? (i64)atomicAdd(&lc->act_n, 1) : (i64)atomicAdd((u64*)&M.state[MS_ACT_N], 1ULL);
            M.act[idx] = (int32_t)s;
        }
    }
    bool enters = T2 > 0 && now >= T2 && now - f <= T2;
    if (enters) {
        const u64 pa_thr = lc ? lc->t2pa : (u64)__ldcg(&M.state[MS_T2_PA]);
        if (pa_thr == 0) enters = now - f < T2;
        else {
            // grid mode, new product token: its prefix is being written by CTA 0 right now -- count the pair in (a superset is safe)
            const u64 pre = (!lc && expect_new) ? ~0ULL : M.tok_pre[x];
            const bool was_in = now - f == T2 && pre >= pa_thr, is_in = now > T2 || pre >= pa_thr;
            enters = is_in && !was_in;
        }
    }
    if (enters) {
        const uint32_t bit = 1u << (s & 31);
        if (!(atomicOr(&M.intop[s >> 5], bit) & bit)) {
            const i64 idx = lc ? (i64)atomicAdd(&lc->top_n, 1) : (i64)atomicAdd((u64*)&M.state[MS_TOP_N], 1ULL);
            if (idx < ML_TOP_N) {
                M.top_slot[idx] = (int32_t)s; M.top_key[idx] = key;
                if (lc) { lc->tslot[idx] = (int32_t)s; lc->tkey[idx] = key; lc->LM.pending[idx] = 1; }
            } else {
                atomicAnd(&M.intop[s >> 5], ~bit);
                if (lc) lc->top_ovf = 1; else M.state[MS_TOP_OVF] = 1;
            }
        }
    }
    return (int32_t)s;
}
// Leader mode, after the rewrite of a merge with a new product token: thresholds of the pairs it created.
// Each new pair is visited exactly once, so the membership bits need no test-and-set.
__device__ __forceinline__ void leader_new_pairs(const MergeParams& M, LeaderCtx* lc, i64 T, i64 T2) {
    const int n = lc->nnew;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int32_t s = i < ML_NEWP_N ? lc->newp[i] : M.newp[i - ML_NEWP_N];
        const u64 key = i < ML_NEWP_N ? lc->newp_key[i] : __ldcg(&M.pkey[s]);
        const int32_t ka = (int32_t)((key >> 32) & 0x7fffffff), kb = (int32_t)(key & 0xffffffffu);
        // one round trip: the count and (speculatively, most new pairs stay below T2) the token data of a top-list entry
        const i64 cnt = __ldcg(&M.pcnt[s]);
        const u64 pre_a = M.tok_pre[ka], pre_b = M.tok_pre[kb], h_a = M.tok_hash[ka], h_b = M.tok_hash[kb], pw_b = M.tok_pow[kb];
        if (cnt < T) continue;
        const uint32_t bit = 1u << (s & 31);
        atomicOr(&M.inact[s >> 5], bit);
        M.act[atomicAdd(&lc->act_n, 1)] = s;
        if (cnt > T2 || (cnt == T2 && pre_a >= lc->t2pa)) {
            const int idx = atomicAdd(&lc->top_n, 1);
            if (idx < ML_TOP_N) {
                atomicOr(&M.intop[s >> 5], bit);
                M.top_slot[idx] = s; M.top_key[idx] = key;
                lc->tslot[idx] = s; lc->tkey[idx] = key; mirror_set(&lc->LM, idx, cnt);
                lc->tpa[idx] = pre_a; lc->tpb[idx] = pre_b; lc->tha[idx] = h_a; lc->thb[idx] = h_b; lc->tpwb[idx] = pw_b;
                lc->tp0[idx] = 0; lc->tplen[idx] = 0;      // created after the last index rebuild: no postings yet
                mirror_insert(&lc->LM, s, idx);
            } else lc->top_ovf = 1;
        }
    }
}
// seg_ctr != nullptr: batched merges -- every member of a batch owns a reserved segment [seg_base, seg_base + its candidate count)
// of the log (the words one merge rewrites must stay contiguous), filled through the member's own counter.
__device__ __forceinline__ void alog_append(const MergeParams& M, int32_t w, i64 off, LeaderCtx* lc, int* seg_ctr = nullptr, i64 seg_base = 0) {
    if (seg_base < 0) return;                  // batched merges: the member's segment is a copy of its candidate list (rewrite_batch)
    const i64 d = seg_ctr ? seg_base + (i64)atomicAdd(seg_ctr, 1)
                          : (lc ? (i64)atomicAdd(&lc->alog_n, 1) : (i64)atomicAdd((u64*)&M.state[MS_ALOG_N], 1ULL));
    if (d < M.alog_cap) M.alog_word[d] = POST_PACK(w, off);
    else atomicOr((u64*)&M.state[MS_ERROR], (u64)ME_INTERNAL);     // callers reserve space up front
}

// one thread rewrites one word in place (left->right, non-overlapping) and applies the deltas
// `cur` = table slot of the pair being merged.  Every site would decrement that ONE counter (thousands of atomics on a single
// address in a heavy merge: they serialise in the L2); the merge removes every occurrence of the pair (trainer.py:268-285:
// its count ends at 0 and the key is deleted), so nobody decrements it and its count is stored as 0 once per merge.
__device__ void rewrite_word_thread(const MergeParams& M, int32_t w, int32_t a, int32_t b, int32_t c, i64 T, i64 T2, LeaderCtx* lm, bool xnew, int32_t cur,
                                    int* seg_ctr = nullptr, i64 seg_base = 0) {
    const i64 off = M.woff[w];
    int32_t* s = M.wsym + off;
    int32_t* ws = M.wslot + off;
    int n = M.wlen[w];
    i64 f = M.wcnt[w];
    int o = 0, j = 0;
    int32_t prev_new = -1;
    bool prev_changed = false, any = false;
    while (j < n) {
        int32_t x = s[j];
        if (j + 1 < n && x == a && s[j + 1] == b) {
            const int32_t sl_ab = ws[j];
            if (o > 0) { if (ws[j - 1] != cur) pair_sub(M, ws[j - 1], f, lm); ws[o - 1] = pair_add(M, prev_new, c, f, T, T2, lm, xnew); }
            (void)sl_ab;
            s[o++] = c; prev_new = c; prev_changed = true; any = true; j += 2;
        } else {
            if (o > 0) {
                const int32_t sl_old = ws[j - 1];
                if (prev_changed) { if (sl_old != cur) pair_sub(M, sl_old, f, lm); ws[o - 1] = pair_add(M, prev_new, x, f, T, T2, lm, xnew); }
                else ws[o - 1] = sl_old;
            }
            s[o++] = x; prev_new = x; prev_changed = false; j += 1;
        }
    }
    if (any) { M.wlen[w] = o; alog_append(M, w, off, lm, seg_ctr, seg_base); }
}

// one warp rewrites one word (a != b): every lane owns one old position per 32-symbol chunk, so the
// pair-count updates of a word are issued in parallel instead of as one dependent chain
__device__ void rewrite_word_warp(const MergeParams& M, int32_t w, int32_t a, int32_t b, int32_t c, i64 T, i64 T2, LeaderCtx* lm, bool xnew) {
    const int lane = threadIdx.x & 31;
    const i64 woff_ = M.woff[w];
    int32_t* s = M.wsym + woff_;
    int32_t* ws = M.wslot + woff_;
    const int n = M.wlen[w];
    const i64 f = M.wcnt[w];
    int out = 0;
    bool any = false;
    int32_t carry_prev = -1;                       // old symbol at position base-1 (may already be overwritten)
    for (int base = 0; base < n; base += 32) {
        const int j = base + lane;
        int32_t xm1 = -1, x0 = -1, x1 = -1, x2 = -1, x3 = -1, ps0 = 0;
        if (j < n) x0 = s[j];
        if (j + 1 < n) ps0 = ws[j];
        if (j + 1 < n) x1 = s[j + 1];
        if (j + 2 < n) x2 = s[j + 2];
        if (j + 3 < n) x3 = s[j + 3];
        xm1 = __shfl_up_sync(0xffffffffu, x0, 1);
        if (lane == 0) xm1 = carry_prev;
        carry_prev = __shfl_sync(0xffffffffu, x0, 31);
        __syncwarp();
        const bool valid = j < n;
        const bool sel0 = valid && x0 == a && x1 == b;          // merge site starts here
        const bool rem0 = valid && xm1 == a && x0 == b;         // second half of a site: removed
        const bool sel1 = x1 == a && x2 == b;
        const bool keep = valid && !rem0;
        const unsigned keepmask = __ballot_sync(0xffffffffu, keep);
        any |= __any_sync(0xffffffffu, sel0);
        // old pair (j, j+1) disappears when either side is part of a site
        if (valid && j + 1 < n && !sel0 && (rem0 || sel1)) pair_sub(M, ps0, f, lm);      // sel0: the merged pair itself (see rewrite_word_thread)
        // new pair starting at kept position j
        int32_t nslot = ps0; bool has_pair = false;
        if (keep) {
            const int jn = sel0 ? j + 2 : j + 1;              // next kept old position
            if (jn < n) {
                has_pair = true;
                const int32_t xn = sel0 ? x2 : x1, xnn = sel0 ? x3 : x2;
                const bool seln = xn == a && xnn == b;
                if (sel0 || seln) nslot = pair_add(M, sel0 ? c : x0, seln ? c : xn, f, T, T2, lm, xnew);
            }
        }
        __syncwarp();
        if (keep) {
            const int ni = out + __popc(keepmask & ((1u << lane) - 1));
            if (ni != j || sel0) s[ni] = sel0 ? c : x0;           // untouched prefix of the word: nothing to store
            if (has_pair && (ni != j || nslot != ps0)) ws[ni] = nslot;
        }
        out += __popc(keepmask);
        __syncwarp();
    }
    if (any && lane == 0) { M.wlen[w] = out; alog_append(M, w, woff_, lm); }
}

// 32 / G words per warp, one per G-lane group (G = 8: words are ~6 symbols on average; G = 4: twice as many candidates per
// pass once the words have shrunk).  Group-local version of rewrite_word_warp; w < 0 marks an idle group.  All 32 lanes
// must call it together.
// Returns the word's length after the rewrite (group-uniform; n itself when nothing changed).
template <int G>
__device__ int rewrite_words_g(const MergeParams& M, int32_t w, i64 off, int n, i64 f,
                                int32_t a, int32_t b, int32_t c, i64 T, i64 T2, LeaderCtx* lm, bool xnew,
                                int* seg_ctr = nullptr, i64 seg_base = 0) {
    const int lane = threadIdx.x & 31, gl = lane & (G - 1), gshift = lane & ~(G - 1);
    constexpr unsigned GM = (1u << G) - 1u;
    int32_t* s = M.wsym + (w >= 0 ? off : 0);
    int32_t* ws = M.wslot + (w >= 0 ? off : 0);
    if (w < 0) n = 0;
    int out = 0;
    bool any = false;
    int32_t carry_prev = -1;
    // the symbols of chunk k+1 are loaded before the atomics of chunk k are issued (software pipelining: a word
    // longer than 8 symbols would otherwise pay the L2 round trips of every chunk one after the other);
    // chunk k only stores to positions < G(k+1), chunk k+1 only loads positions >= G(k+1)
    int32_t nx0 = -1, nx1 = -1, nx2 = -1, nx3 = -1, nps0 = 0;
    if (gl < n) nx0 = s[gl];
    if (gl + 1 < n) { nx1 = s[gl + 1]; nps0 = ws[gl]; }
    if (gl + 2 < n) nx2 = s[gl + 2];
    if (gl + 3 < n) nx3 = s[gl + 3];
    for (int base = 0; __any_sync(0xffffffffu, base < n); base += G) {
        const int j = base + gl;
        const int32_t x0 = nx0, x1 = nx1, x2 = nx2, x3 = nx3, ps0 = nps0;
        nx0 = -1; nx1 = -1; nx2 = -1; nx3 = -1; nps0 = 0;
        if (j + G < n) nx0 = s[j + G];
        if (j + G + 1 < n) { nx1 = s[j + G + 1]; nps0 = ws[j + G]; }
        if (j + G + 2 < n) nx2 = s[j + G + 2];
        if (j + G + 3 < n) nx3 = s[j + G + 3];
        int32_t xm1 = __shfl_up_sync(0xffffffffu, x0, 1, G);
        if (gl == 0) xm1 = carry_prev;
        carry_prev = __shfl_sync(0xffffffffu, x0, G - 1, G);
        __syncwarp();
        if (base == 0) ML_TRG(13);
        const bool valid = j < n;
        const bool sel0 = valid && x0 == a && x1 == b;
        const bool rem0 = valid && xm1 == a && x0 == b;
        const bool sel1 = x1 == a && x2 == b;
        const bool keep = valid && !rem0;
        const unsigned keepmask = (__ballot_sync(0xffffffffu, keep) >> gshift) & GM;
        any |= ((__ballot_sync(0xffffffffu, sel0) >> gshift) & GM) != 0;
        if (valid && j + 1 < n && !sel0 && (rem0 || sel1)) pair_sub(M, ps0, f, lm);      // sel0: the merged pair itself (see rewrite_word_thread)
        if (base == 0) ML_TRG(14);
        int32_t nslot = ps0; bool has_pair = false;
        if (keep) {
            const int jn = sel0 ? j + 2 : j + 1;
            if (jn < n) {
                has_pair = true;
                const int32_t xn = sel0 ? x2 : x1, xnn = sel0 ? x3 : x2;
                const bool seln = xn == a && xnn == b;
                if (sel0 || seln) nslot = pair_add(M, sel0 ? c : x0, seln ? c : xn, f, T, T2, lm, xnew);
            }
        }
        __syncwarp();
        if (base == 0) ML_TRG(15);
        if (keep) {
            const int ni = out + __popc(keepmask & ((1u << gl) - 1));
            if (ni != j || sel0) s[ni] = sel0 ? c : x0;           // untouched prefix of the word: nothing to store
            if (has_pair && (ni != j || nslot != ps0)) ws[ni] = nslot;
        }
        out += __popc(keepmask);
        __syncwarp();
    }
    if (any && gl == 0) { M.wlen[w] = out; alog_append(M, w, off, lm, seg_ctr, seg_base); }
    return any ? out : n;
}

// candidate ranges for pair (a, b) at slot: CSR postings + affected-log segments of the merges that
// produced a or b since the last index rebuild.  Returns the number of ranges, or -1 if too many.
struct Ranges { const i64* base[ML_MAX_RANGES]; int len[ML_MAX_RANGES]; int n; i64 total; };

__device__ void build_ranges(const MergeParams& M, int32_t slot, int32_t a, int32_t b, Ranges* R, bool have_post = false, uint32_t post0 = 0, uint32_t postlen = 0) {
    int n = 0; i64 total = 0;
    // the four loads below are independent: one round trip in the common case (each token made by <= 1 merge since the rebuild);
    // the leader passes the postings it cached with the top-list entry (have_post), leaving only the L2-resident token heads
    const uint32_t p0 = have_post ? post0 : M.ioff[slot], p1 = have_post ? post0 + postlen : M.ioff[slot + 1];
    const int4 ha = M.tok_head[a], hb = M.tok_head[b];
    if (p1 > p0) { R->base[n] = M.ipost + p0; R->len[n] = (int)(p1 - p0); total += p1 - p0; n++; }
    for (int side = 0; side < 2; side++) {
        if (side == 1 && b == a) break;
        const int4 h = side == 0 ? ha : hb;
        if (h.x < 0) continue;
        if (h.z > h.y) {
            if (n >= ML_MAX_RANGES) { R->n = -1; R->total = total; return; }
            R->base[n] = M.alog_word + h.y; R->len[n] = h.z - h.y; total += h.z - h.y; n++;
        }
        for (int32_t mm = h.w; mm >= 0; mm = M.merge_next[mm]) {
            int ln = M.seg_end[mm] - M.seg_start[mm];
            if (ln <= 0) continue;
            if (n >= ML_MAX_RANGES) { R->n = -1; R->total = total; return; }
            R->base[n] = M.alog_word + M.seg_start[mm]; R->len[n] = ln; total += ln; n++;
        }
    }
    R->n = n; R->total = total;
}
__device__ __forceinline__ i64 range_item(const Ranges& R, i64 it) {
    for (int r = 0; r < R.n; r++) {
        if (it < R.len[r]) { const i64* b = R.base[r]; __builtin_assume(__isGlobal(b)); return b[it]; }    // not a generic load
        it -= R.len[r];
    }
    return -1;
}

// merged token id for (a, b): existing id when the bytes are already a token, else n_tok
struct MergedInfo { u64 H, P; i64 oa, ob, la, lb, tslot; };

__device__ int32_t lookup_merged(const MergeParams& M, int32_t a, int32_t b, int32_t n_tok, MergedInfo* info = nullptr) {
    const u64 ha = M.tok_hash[a], hb = M.tok_hash[b], pa_ = M.tok_pow[a], pb_ = M.tok_pow[b];
    u64 H = ha * pb_ + hb;
    i64 oa = M.tok_off[a], ob = M.tok_off[b];
    i64 la = M.tok_off[a + 1] - oa, lb = M.tok_off[b + 1] - ob;
    u64 mask = (u64)M.tset_cap - 1, slot = mix64(H) & mask;
    if (info) { info->H = H; info->P = pa_ * pb_; info->oa = oa; info->ob = ob; info->la = la; info->lb = lb; }
    for (;;) {
        u64 e = *(volatile u64*)&M.tset[slot];
        if (e == 0) { if (info) info->tslot = (i64)slot; return n_tok; }
        int32_t id = (int32_t)(e & 0xffffffffu) - 1;
        if ((e >> 32) == (H >> 32) && id < n_tok && M.tok_hash[id] == H && M.tok_off[id + 1] - M.tok_off[id] == la + lb) {
            const uint8_t* pc = M.tok_bytes + M.tok_off[id];
            const uint8_t* pa = M.tok_bytes + oa;
            const uint8_t* pb = M.tok_bytes + ob;
            bool eq = true;
            for (i64 k = 0; k < la && eq; k++) eq = pc[k] == pa[k];
            for (i64 k = 0; k < lb && eq; k++) eq = pc[la + k] == pb[k];
            if (eq) return id;
        }
        slot = (slot + 1) & mask;
    }
}

// leader mode: the hash of the merged bytes comes from the top-list entry (shared memory), so the token-set probe and
// the loads the commit needs (offsets, base^len(a)) are ONE round trip instead of two
__device__ int32_t lookup_merged_leader(const MergeParams& M, const LeaderCtx& C, int idx, int32_t a, int32_t b, int32_t n_tok, MergedInfo* info) {
    const u64 H = C.tha[idx] * C.tpwb[idx] + C.thb[idx];
    const u64 mask = (u64)M.tset_cap - 1;
    u64 slot = mix64(H) & mask;
    u64 e = *(volatile u64*)&M.tset[slot];
    const i64 oa = M.tok_off[a], oa1 = M.tok_off[a + 1], ob = M.tok_off[b], ob1 = M.tok_off[b + 1];
    const u64 pa_ = M.tok_pow[a];
    const i64 la = oa1 - oa, lb = ob1 - ob;
    info->H = H; info->P = pa_ * C.tpwb[idx]; info->oa = oa; info->ob = ob; info->la = la; info->lb = lb;
    for (;;) {
        if (e == 0) { info->tslot = (i64)slot; return n_tok; }
        int32_t id = (int32_t)(e & 0xffffffffu) - 1;
        if ((e >> 32) == (H >> 32) && id < n_tok && M.tok_hash[id] == H && M.tok_off[id + 1] - M.tok_off[id] == la + lb) {
            const uint8_t* pc = M.tok_bytes + M.tok_off[id];
            const uint8_t* pa = M.tok_bytes + oa;
            const uint8_t* pb = M.tok_bytes + ob;
            bool eq = true;
            for (i64 k = 0; k < la && eq; k++) eq = pc[k] == pa[k];
            for (i64 k = 0; k < lb && eq; k++) eq = pc[la + k] == pb[k];
            if (eq) return id;
        }
        slot = (slot + 1) & mask;
        e = *(volatile u64*)&M.tset[slot];
    }
}

// record merge m and (when the bytes are new) create token c; executed by ONE block
__device__ void commit_merge(const MergeParams& M, i64 m, int32_t a, int32_t b, int32_t c, bool is_new, i64 alog_start) {
    if (threadIdx.x == 0) {
        M.merges[2 * m] = a; M.merges[2 * m + 1] = b; M.merge_new[m] = c;
        M.seg_start[m] = (int32_t)alog_start;
        M.merge_next[m] = M.tok_first[c];
    }
    if (is_new) {
        i64 oa = M.tok_off[a], ob = M.tok_off[b], oc = M.tok_off[c];
        i64 la = M.tok_off[a + 1] - oa, lb = M.tok_off[b + 1] - ob;
        if (oc + la + lb > M.tok_bytes_cap || c + 1 >= M.max_tokens) {
            if (threadIdx.x == 0) atomicOr((u64*)&M.state[MS_ERROR], (u64)ME_TOK_POOL_FULL);
        } else {
            for (i64 k = threadIdx.x; k < la; k += blockDim.x) M.tok_bytes[oc + k] = M.tok_bytes[oa + k];
            for (i64 k = threadIdx.x; k < lb; k += blockDim.x) M.tok_bytes[oc + la + k] = M.tok_bytes[ob + k];
            __syncthreads();
            if (threadIdx.x == 0) {
                u64 H = M.tok_hash[a] * M.tok_pow[b] + M.tok_hash[b];
                M.tok_off[c + 1] = oc + la + lb;
                M.tok_hash[c] = H; M.tok_pow[c] = M.tok_pow[a] * M.tok_pow[b];
                M.tok_pre[c] = tok_prefix_concat(M.tok_pre[a], la, M.tok_pre[b]);
                M.tok_first[c] = -1; M.tok_head[c].x = -1;
                __threadfence();
                u64 mask = (u64)M.tset_cap - 1, slot = mix64(H) & mask;
                while (M.tset[slot] != 0) slot = (slot + 1) & mask;
                atomicExch(&M.tset[slot], (H & 0xffffffff00000000ULL) | (u64)(uint32_t)(c + 1));
                M.state[MS_NTOK] = c + 1; M.state[MS_POOL_USED] = oc + la + lb;
            }
        }
    }
}
// leader-mode commit by one warp; MI comes from lookup_merged, oc = current end of the token byte pool
__device__ void commit_merge_leader(const MergeParams& M, i64 m, int32_t a, int32_t b, int32_t c, bool is_new, i64 alog_start,
                                    const MergedInfo& MI, i64 oc) {
    const int lane = threadIdx.x & 31;
    if (lane == 0) {
        M.merges[2 * m] = a; M.merges[2 * m + 1] = b; M.merge_new[m] = c;
        M.seg_start[m] = (int32_t)alog_start;
        M.merge_next[m] = is_new ? -1 : M.tok_first[c];
    }
    if (!is_new) return;
    if (oc + MI.la + MI.lb > M.tok_bytes_cap || c + 1 >= M.max_tokens) {
        if (lane == 0) atomicOr((u64*)&M.state[MS_ERROR], (u64)ME_TOK_POOL_FULL);
        return;
    }
    for (i64 k = lane; k < MI.la; k += 32) M.tok_bytes[oc + k] = M.tok_bytes[MI.oa + k];
    for (i64 k = lane; k < MI.lb; k += 32) M.tok_bytes[oc + MI.la + k] = M.tok_bytes[MI.ob + k];
    if (lane == 0) {
        M.tok_off[c + 1] = oc + MI.la + MI.lb;
        M.tok_hash[c] = MI.H; M.tok_pow[c] = MI.P;
        M.tok_pre[c] = tok_prefix_concat(M.tok_pre[a], MI.la, M.tok_pre[b]);
        M.tok_first[c] = -1; M.tok_head[c].x = -1;
        *(volatile u64*)&M.tset[MI.tslot] = (MI.H & 0xffffffff00000000ULL) | (u64)(uint32_t)(c + 1);
        M.state[MS_NTOK] = c + 1; M.state[MS_POOL_USED] = oc + MI.la + MI.lb;
    }
}
// after the rewrite of merge m finished: close its segment and link it to its product token
__device__ __forceinline__ void close_merge(const MergeParams& M, i64 m, int32_t c) {
    const int32_t e = (int32_t)__ldcg(&M.state[MS_ALOG_N]);
    M.seg_end[m] = e;
    M.tok_first[c] = (int32_t)m;
    M.tok_head[c] = make_int4((int32_t)m, M.seg_start[m], e, M.merge_next[m]);
    M.state[MS_NMERGES] = m + 1;
}

// ---- leader mode: CTA 0 runs merges alone, block barriers only ---------------------------------

// best entry of the top list, computed by ONE block (every block gets the same answer when all call it)
__device__ Best top_best(const MergeParams& M, i64 top_n, Best* sh_best, i64* sh_cnt) {
    Best mine{0, -1, 0, 0, 0};
    const int tn = top_n < ML_TOP_N ? (int)top_n : ML_TOP_N;
    if ((int)threadIdx.x < tn) {
        const int32_t sl = M.top_slot[threadIdx.x];
        const u64 k = M.top_key[threadIdx.x];
        const i64 cnt = __ldcg(&M.pcnt[sl]);
        if (cnt > 0) mine = Best{cnt, sl, (int32_t)((k >> 32) & 0x7fffffff), (int32_t)(k & 0xffffffffu), 0};
    }
    return block_best_fast(M, mine, sh_best, sh_cnt);
}

#define LR_TOP 1      // leader stopped: top list exhausted / overflowed
#define LR_OTHER 2    // leader stopped: the next merge needs the whole grid (or nothing is left to do)
#define LR_REBUILD 3  // leader stopped: time to fold the affected-word log into the CSR index (fewer stale candidates)


// ---- batch selection -------------------------------------------------------------------------------
// The ML_SEL = ML_BATCH_MAX + 1 largest entries of the top list in strictly descending (count, then list index) order.
// key = count << 9 | (511 - index); 0 = none.  Phase 1: every warp that holds list entries extracts the ML_SEL largest keys
// of its 32 lanes (redux.sync); phase 2 (after ONE block barrier): every warp redundantly extracts the ML_SEL largest of the
// per-warp lists, so all warps hold the same result without a second barrier.  Returned per lane: lane r gets the r-th key.
#ifndef ML_BATCH_MAX
#define ML_BATCH_MAX 31       // one lane per member, and one more lane for the first entry left out
#endif
#define ML_SEL (ML_BATCH_MAX + 1)          // entries put in exact order by the selection (<= 32: one per lane of a warp)
#define ML_HEAD 64                         // capacity of the head list the selection starts from
#define ML_HSEL 5                          // entries the prefetch helpers look at (select_top)
#define ML_SEL_WARPS (ML_TOP_N / 32)
__device__ __forceinline__ u64 warp_max_u64(u64 v) {
    const uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
    const uint32_t mhi = __reduce_max_sync(0xffffffffu, hi);
    const uint32_t mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
    return ((u64)mhi << 32) | mlo;
}
__device__ __forceinline__ u64 select_top(u64 key, u64* sh_keys /* [ML_SEL_WARPS][ML_HSEL] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp < ML_SEL_WARPS) {
        u64 k = key;
#pragma unroll
        for (int r = 0; r < ML_HSEL; r++) {
            const u64 m = warp_max_u64(k);
            if (k == m) k = 0;                        // keys are unique (the index is part of the key)
            if (lane == r) sh_keys[warp * ML_HSEL + r] = m;
        }
    }
    __syncthreads();
    constexpr int NV = (ML_SEL_WARPS * ML_HSEL + 31) / 32;
    u64 v[NV];
#pragma unroll
    for (int u = 0; u < NV; u++) { const int i = lane + 32 * u; v[u] = i < ML_SEL_WARPS * ML_HSEL ? sh_keys[i] : 0; }
    u64 mine = 0;
#pragma unroll
    for (int r = 0; r < ML_HSEL; r++) {
        u64 lm = v[0];
#pragma unroll
        for (int u = 1; u < NV; u++) lm = v[u] > lm ? v[u] : lm;
        const u64 m = warp_max_u64(lm);
        if (lane == r) mine = m;
#pragma unroll
        for (int u = 0; u < NV; u++) if (v[u] == m) v[u] = 0;
    }
    return mine;
}

// Stage C of the leader: every G-lane group of the first nwarps-1 warps takes candidates it0, it0 + ngroups, ...
// Software pipeline over the passes of a group: the candidate of pass p+2 and the word header of pass p+1 are loaded
// while pass p is rewritten, so only the first pass pays those round trips in full (the word arrays of a large corpus
// do not fit the L2: every round trip is a DRAM access).
template <int G>
__device__ __forceinline__ void leader_rewrite(const MergeParams& M, LeaderCtx& C, const Ranges& R, int warp, int lane, int nwarps,
                                               int32_t a, int32_t b, int32_t c, i64 T, i64 T2, bool is_new, int32_t cur) {
    constexpr int GPW = 32 / G;                      // groups per warp
    const int ngroups = (nwarps - 1) * GPW, gl = lane & (G - 1), lead = lane & ~(G - 1);
    const int total = (int)R.total;
    const int it0 = warp * GPW + lane / G;
    const i64 e_cur = it0 < total ? range_item(R, it0) : -1, e_nx = it0 + ngroups < total ? range_item(R, it0 + ngroups) : -1;
    int32_t w_cur = POST_WORD(e_cur), w_nx = POST_WORD(e_nx);
    i64 o_nx = POST_OFF(e_nx);
    {
        int take = (w_cur >= 0 && gl == 0) ? (dedupe_claim(&C, w_cur) ? 1 : 0) : 0;
        take = __shfl_sync(0xffffffffu, take, lead);
        if (!take) w_cur = -1;
    }
    uint32_t off_cur = 0; int n_cur = 0; i64 f_cur = 0;
    if (w_cur >= 0) {                       // header and symbols in ONE round trip: the entry carries the symbol slot
        off_cur = (uint32_t)POST_OFF(e_cur);
        if (gl == 0) { prefetch_l2(&M.wsym[off_cur]); prefetch_l2(&M.wslot[off_cur]); }
        n_cur = M.wlen[w_cur]; f_cur = M.wcnt[w_cur];
    }
    for (int base = 0; base + warp * GPW < total; base += ngroups) {     // warps without a candidate go straight to the barrier
        const int it2 = base + it0 + 2 * ngroups;
        const i64 e_nx2 = it2 < total ? range_item(R, it2) : -1;
        {
            int take = (w_nx >= 0 && gl == 0) ? (dedupe_claim(&C, w_nx) ? 1 : 0) : 0;
            take = __shfl_sync(0xffffffffu, take, lead);
            if (!take) w_nx = -1;
        }
        uint32_t off_nx = 0; int n_nx = 0; i64 f_nx = 0;
        if (w_nx >= 0) {
            off_nx = (uint32_t)o_nx;
            if (gl == 0) { prefetch_l2(&M.wsym[off_nx]); prefetch_l2(&M.wslot[off_nx]); }
            n_nx = M.wlen[w_nx]; f_nx = M.wcnt[w_nx];
        }
        if (a != b) rewrite_words_g<G>(M, w_cur, (i64)off_cur, n_cur, f_cur, a, b, c, T, T2, &C, is_new);
        else if (w_cur >= 0 && gl == 0) rewrite_word_thread(M, w_cur, a, b, c, T, T2, &C, is_new, cur);
        w_cur = w_nx; off_cur = off_nx; n_cur = n_nx; f_cur = f_nx; w_nx = POST_WORD(e_nx2); o_nx = POST_OFF(e_nx2);
    }
}

// warp-wide maximum of non-negative 64-bit values with two redux.sync (32-bit) instead of ten shuffles
__device__ __forceinline__ i64 warp_max_i64(i64 v) {
    const uint32_t hi = (uint32_t)((u64)v >> 32), lo = (uint32_t)(u64)v;
    const uint32_t mhi = __reduce_max_sync(0xffffffffu, hi);
    const uint32_t mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
    return (i64)(((u64)mhi << 32) | mlo);
}

// block-wide maximum of a 64-bit value, the same result in every thread (two block barriers; sh: one word per warp)
__device__ __forceinline__ u64 block_max_u64(u64 v, u64* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const u64 wm = warp_max_u64(v);
    __syncthreads();                       // sh may still be read from the previous use
    if (lane == 0) sh[warp] = wm;
    __syncthreads();
    return warp_max_u64(lane < nwarps ? sh[lane] : 0ULL);
}

// ---- stage A of the leader, one merge: best entry of the whole top list --------------------------------------------------
// `mine` = this thread's entry (pad = list index).  Max count first (redux), ties decided by the cached 8-byte prefixes of the
// tokens and, where those are equal, by the bytes.  Block-wide: every thread returns the same entry.
__device__ Best leader_argmax_one(const MergeParams& M, LeaderCtx& C, const Best& mine, i64* sh_wmax, int* sh_ncand, Best* sh_cand, Best* sh_best) {
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, lane = threadIdx.x & 31;
    {
        const i64 wm = warp_max_i64(mine.cnt);
        if (lane == 0) sh_wmax[warp] = wm;
    }
    __syncthreads();
    const i64 mx = warp_max_i64(lane < nwarps ? sh_wmax[lane] : 0);
    if (mine.slot >= 0 && mine.cnt == mx) { const int ci = atomicAdd(sh_ncand, 1); if (ci < 32) sh_cand[ci] = mine; }
    __syncthreads();
    Best best{0, -1, 0, 0, 0};
    const int ncand = *sh_ncand;
    if (mx <= 0) return best;
    if (ncand == 1) return sh_cand[0];
    if (ncand <= 32) {                          // ties: byte-wise comparison, every warp redundantly (no barrier)
        Best t = lane < ncand ? sh_cand[lane] : Best{0, -1, 0, 0, 0};
        // fast path: lexicographic maximum of (prefix of a, prefix of b) with four redux.sync rounds.  Different
        // prefixes order the tokens; equal prefixes decide only when the tokens themselves are equal.
        {
            bool alive = lane < ncand;
            const u64 pa = alive ? C.tpa[t.pad] : 0ULL, pb = alive ? C.tpb[t.pad] : 0ULL;
            uint32_t wv = (uint32_t)(pa >> 32), mw = __reduce_max_sync(0xffffffffu, alive ? wv : 0u); alive = alive && wv == mw;
            wv = (uint32_t)pa; mw = __reduce_max_sync(0xffffffffu, alive ? wv : 0u); alive = alive && wv == mw;
            const uint32_t amax = __reduce_max_sync(0xffffffffu, alive ? (uint32_t)t.a : 0u);
            const uint32_t amin = __reduce_min_sync(0xffffffffu, alive ? (uint32_t)t.a : 0xffffffffu);
            if (amax == amin) {                                  // every survivor has the SAME left token: the right one decides
                wv = (uint32_t)(pb >> 32); mw = __reduce_max_sync(0xffffffffu, alive ? wv : 0u); alive = alive && wv == mw;
                wv = (uint32_t)pb; mw = __reduce_max_sync(0xffffffffu, alive ? wv : 0u); alive = alive && wv == mw;
                const uint32_t am = __ballot_sync(0xffffffffu, alive);
                if (__popc(am) == 1) return shfl_best_from(t, __ffs(am) - 1);
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            Best u = shfl_best(t, o);
            u.pad = __shfl_xor_sync(0xffffffffu, t.pad, o);
            bool gt = false;                                      // equal counts: (left bytes, right bytes)
            if (u.slot >= 0 && t.slot < 0) gt = true;
            else if (u.slot >= 0 && u.slot != t.slot) {
                int r = tok_cmp_pre(M, u.a, C.tpa[u.pad], t.a, C.tpa[t.pad]);
                if (r == 0) r = tok_cmp_pre(M, u.b, C.tpb[u.pad], t.b, C.tpb[t.pad]);
                gt = r > 0;
            }
            if (gt) t = u;
        }
        return t;
    }
    // More than 32 pairs share the maximum (the tie regime: up to the whole list).  Byte-wise comparisons across the
    // block cost thousands of cycles; the cached prefixes decide almost every time: the greatest LEFT prefix, then --
    // if every survivor has the same left token -- the greatest RIGHT prefix.  Only what is still tied goes through
    // the byte-wise reduction.
    __syncthreads();
    bool alive = mine.slot >= 0 && mine.cnt == mx;
    const u64 mpa = block_max_u64(alive ? C.tpa[threadIdx.x] : 0ULL, (u64*)sh_wmax);
    alive = alive && C.tpa[threadIdx.x] == mpa;
    const uint32_t amax = (uint32_t)block_max_u64(alive ? (u64)(uint32_t)mine.a : 0ULL, (u64*)sh_wmax);
    const uint32_t namin = (uint32_t)block_max_u64(alive ? (u64)(0xffffffffu - (uint32_t)mine.a) : 0ULL, (u64*)sh_wmax);
    if (amax == 0xffffffffu - namin) {            // one left token: the right one decides
        const u64 mpb = block_max_u64(alive ? C.tpb[threadIdx.x] : 0ULL, (u64*)sh_wmax);
        alive = alive && C.tpb[threadIdx.x] == mpb;
    }
    const int left = __syncthreads_count(alive);
    if (left == 1) {
        if (alive) sh_cand[0] = mine;
        __syncthreads();
        return sh_cand[0];
    }
    best = block_best(M, alive ? mine : Best{0, -1, 0, 0, 0}, sh_best);
    if (best.slot >= 0) best.pad = mirror_find(&C.LM, best.slot);      // block_best does not carry the list index
    return best;
}

// The head of the top list: theta such that between 5/8 head_cap and head_cap entries have a count >= theta, found with a 32-bin histogram of
// the counts in [lo, max] that zooms into the bin that did not fit (counts repeat: a single count can hold more entries than
// fit, then fewer -- possibly none -- are taken and *sticky tells the caller not to ask again until the head is empty).
// Block-wide, every thread returns the same value; ends with a barrier.
// hi0 > 0: the caller knows that exactly above0 entries have a count >= hi0 (the old head, nearly used up): the search starts
// below it and the block-wide maximum is not needed.
__device__ i64 leader_head_threshold(const Best& mine, i64 lo, i64* sh_wmax, int* hist /* [34] */, bool* sticky, int head_cap, i64 hi0 = 0, int above0 = 0) {
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, lane = threadIdx.x & 31;
    i64 hi = hi0;
    int above = above0;
    if (hi0 <= 0) {
        const i64 wm = warp_max_i64(mine.slot >= 0 ? mine.cnt : 0);
        __syncthreads();
        if (lane == 0) sh_wmax[warp] = wm;
        __syncthreads();
        hi = warp_max_i64(lane < nwarps ? sh_wmax[lane] : 0) + 1;          // no entry has a count >= hi
        above = 0;
    }
    i64 theta = lo;
    if (hi <= lo) { *sticky = true; __syncthreads(); return lo; }           // nothing above lo at all
    for (int round = 0; round < 64; round++) {
        if (threadIdx.x < 32) hist[threadIdx.x] = 0;
        __syncthreads();
        const i64 width = (hi - lo + 31) / 32;
        if (mine.slot >= 0 && mine.cnt >= lo && mine.cnt < hi) atomicAdd(&hist[(int)((mine.cnt - lo) / width)], 1);
        __syncthreads();
        if (warp == 0) {                               // bins from the top while they fit: suffix sums by shuffles, one warp
            int suf = hist[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_down_sync(0xffffffffu, suf, o); if (lane + o < 32) suf += t; }
            const uint32_t fits = __ballot_sync(0xffffffffu, above + suf <= head_cap);  // an upper range of bins (suf falls with the bin)
            const int bl = fits ? __ffs(fits) - 1 : 32;
            const int cumw = above + (bl < 32 ? __shfl_sync(0xffffffffu, suf, bl & 31) : 0);
            if (lane == 0) { hist[32] = bl - 1; hist[33] = cumw; }
        }
        __syncthreads();
        const int b = hist[32], cum = hist[33];
        theta = lo + (i64)(b + 1) * width;
        __syncthreads();                                                   // everybody has read the result
        if (b < 0) { theta = lo; above = cum; break; }
        if (cum >= head_cap / 2 + head_cap / 8 || width == 1) { above = cum; break; }
        above = cum; hi = theta; lo = lo + (i64)b * width;
    }
    *sticky = above < head_cap / 4;
    return theta;
}

// ---- batched leader merges ------------------------------------------------------------------------------------------
// The reference merges ONE pair per iteration (trainer.py:241-300).  Most consecutive merges do not interact, and the
// leader's cost per iteration is a chain of dependent round trips that does not depend on how much it rewrites -- so the
// leader takes the k best pairs p_1 .. p_k of the top list in ONE iteration whenever it can PROVE that the sequential loop
// would pick exactly these, in this order, with exactly the same state after the k-th:
//   (1) p_1 .. p_k are the first k pairs in the exact order (count, left bytes, right bytes); every pair left out with the
//       count of p_k comes after it in that order and touches no member (2); c_k >= T2: all pairs with a count >= c_k are on
//       the top list (it is complete from T2 up while the threshold is a plain count), and the selection has seen them all;
//   (2) no member "touches" an earlier one: merging (a_i, b_i) only decrements pairs (x, a_i) and (b_i, y) and only creates
//       pairs that contain the new token, so for i < j:  b_j != a_i and a_j != b_i  keep c_j unchanged; every pair created
//       by the batch is bounded by the count of an OLD pair (x, a_i) or (b_i, y) != p_i, which is not a member (2) and
//       therefore < c_k (1: a left-out pair with the count c_k does not touch): nothing new can overtake a member.
//       (a_i == b_i breaks that bound -- "aaaa" makes (aa, aa) out of (a, a) itself -- so such a pair is only taken as
//       the LAST member.)
//   (3) the merged bytes of a member are a NEW token (trainer.py:296-300 gives no new id otherwise, and pairs with an
//       existing token can GAIN above c_k): a member whose bytes exist already is the last one; two members with the
//       same merged bytes never share a batch.
// Old pairs only lose during a batch, so the bounds hold throughout.  Words are rewritten word by word: whoever claims a
// candidate word applies members 1 .. k to it in order -- the state of a word depends on nothing but the word, and the
// pair counts are sums over words, so the result is the sequential one bit for bit.  More equal counts at the top than the
// selection holds, the tie regime (T2pa != 0) and single heavy merges take the one-merge path below.  Grid mode batches by the
// same rules (k_merge_loop): every CTA selects the same members and the groups of all CTAs share the items.
struct BatchSel {
    u64 S[ML_HEAD]; int nS;                     // head of the top list: keys (count << 9 | 511 - index) of the entries with count >= theta
    u64 Skey[ML_HEAD]; int32_t Sslot[ML_HEAD];  // and their pair keys / table slots (by position in S)
    u64 Q[ML_SEL], Qkey[ML_SEL]; int32_t Qslot[ML_SEL];      // the ML_SEL best of the head by (count, list index)
    int hist[34];
    u64 Spa[ML_SEL], Spb[ML_SEL];               // token prefixes of the selected entries (filled when equal counts need ordering)
    struct { int32_t a, b, slot, idx; i64 cnt; } mem[ML_SEL];      // the selected entries in exact order; the first nb are the batch
    int nb;
};
struct BatchCtx {
    Ranges R[ML_BATCH_MAX];
    MergedInfo MI[ML_BATCH_MAX];
    int32_t c[ML_BATCH_MAX];
    u64 mkey[ML_BATCH_MAX];                     // the members' pair keys (the rewrite tests every adjacency of a word against them)
    BatchSel sel;
};

// ---- stage A, batch: which of the best entries of the top list may be merged together ---------------------------------------
// `mine` = this thread's entry of the top list (threads >= tn: none; pad = list index), with its current count.  The entries with
// count >= theta (the HEAD, at most 32) put themselves on a list; warp 0 takes the ML_SEL best of them in exact order (count, then
// left bytes, right bytes) and decides how many form a batch (rules (1) - (2) above; rule (3) needs the token lookups and is
// applied by the caller).  theta / sticky persist between calls; theta is re-chosen when the head runs empty or overflows;
// head_cap (<= ML_HEAD) is the size the head is refilled to: 32 for the leader's small batches, 64 for grid mode's.
// tpa / tpb: 8-byte prefixes of the entries' tokens by list index (leader: shared memory), or nullptr: M.tok_pre is read.
// Block-wide; every thread returns the same count nb (0: the caller takes the one-merge path) and finds the members in BS.mem.
// Deterministic in its inputs: CTAs that see the same list and counts (grid mode) choose the same batch.
__device__ __forceinline__ int select_batch(const MergeParams& M, BatchSel& BS, const Best& mine, int tn, int batch_max, i64 T, i64 Tmin, i64 T2,
                            const u64* tpa, const u64* tpb, i64* sh_wmax, i64& theta, bool& theta_sticky, int head_cap) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __builtin_assume(__isShared(&BS));
    const u64 mykey = mine.slot >= 0 ? (((u64)mine.cnt << 9) | (u64)(511 - (int)threadIdx.x)) : 0ULL;
    int nS;
#if ML_RW_TRACE
    long long _s0 = clock64();
#endif
    for (int pass = 0;; pass++) {
        if (theta > 0 && mine.slot >= 0 && mine.cnt >= theta) {
            const int p = atomicAdd(&BS.nS, 1);
            if (p < head_cap) { BS.S[p] = mykey; BS.Skey[p] = PAIR_KEY(mine.a, mine.b); BS.Sslot[p] = mine.slot; }
        }
        __syncthreads();
        nS = BS.nS;
        if (pass == 0) SELT(0);
        const bool refresh = theta <= 0 || nS > head_cap || (nS < head_cap / 4 && tn > head_cap && !theta_sticky && theta > 1);
        if (!refresh || pass == 2) break;
        const bool below = theta > (T2 > 1 ? T2 : 1) && nS <= head_cap;                                     // the old head is nearly used up: look below it
        theta = leader_head_threshold(mine, T2 > 1 ? T2 : 1, sh_wmax, BS.hist, &theta_sticky, head_cap, below ? theta : 0, below ? nS : 0);     // block-wide; ends with a barrier
        if (tn <= head_cap) { theta = 1; theta_sticky = false; }
        if (threadIdx.x == 0) BS.nS = 0;
        __syncthreads();
    }
    if (nS > head_cap) nS = 0;                   // cannot happen after a refresh; be safe: the one-merge path decides
    SELT(1);
    if (warp == 0) {
        // One warp, and every step reads what it needs from shared memory with broadcast loads (independent, pipelined): a
        // first version passed the entries around with ~230 dependent shuffles and spent 3 us per call on their latency.
        int32_t ma = 0, mb = 0, mslot = -1, midx = -1; i64 mcnt = 0;
        // rank of every head entry by (count, list index): two entries per lane, every lane walks the whole list
        const u64 k0 = lane < nS ? BS.S[lane] : 0ULL, k1 = lane + 32 < nS ? BS.S[lane + 32] : 0ULL;
        int r0 = 0, r1 = 0;
#pragma unroll 4
        for (int j = 0; j < nS; j++) { const u64 v = BS.S[j]; r0 += v > k0 ? 1 : 0; r1 += v > k1 ? 1 : 0; }
        if (k0 != 0 && r0 < ML_SEL) { BS.Q[r0] = k0; BS.Qkey[r0] = BS.Skey[lane]; BS.Qslot[r0] = BS.Sslot[lane]; }            // keys are unique:
        if (k1 != 0 && r1 < ML_SEL) { BS.Q[r1] = k1; BS.Qkey[r1] = BS.Skey[lane + 32]; BS.Qslot[r1] = BS.Sslot[lane + 32]; }  // the ML_SEL best, in order
        __syncwarp();
        const int nsel = nS < ML_SEL ? nS : ML_SEL;
        u64 sel = lane < nsel ? BS.Q[lane] : 0ULL;
        if (sel != 0) {
            midx = 511 - (int)(sel & 511);
            const u64 kk2 = BS.Qkey[lane];
            ma = (int32_t)((kk2 >> 32) & 0x7fffffff); mb = (int32_t)(kk2 & 0xffffffffu);
            mslot = BS.Qslot[lane]; mcnt = (i64)(sel >> 9);
        }
        SELT(2);
        // entries outside the selection: below theta, or (more than ML_SEL in the head) not above the last selected count
        const i64 g = nS > ML_SEL ? (i64)(BS.Q[ML_SEL - 1] >> 9) : theta - 1;
        // equal counts among the selected: (left bytes, right bytes) order them (exact: prefixes, then the bytes)
        int pos = lane;
        const i64 cdown = lane + 1 < nsel ? (i64)(BS.Q[lane + 1] >> 9) : 0;
        if (__ballot_sync(0xffffffffu, sel != 0 && mcnt == cdown && mcnt > g)) {
            u64 pa = 0, pb = 0;
            if (sel != 0) { pa = tpa ? tpa[midx] : __ldcg(&M.tok_pre[ma]); pb = tpb ? tpb[midx] : __ldcg(&M.tok_pre[mb]); BS.Spa[lane] = pa; BS.Spb[lane] = pb; }
            __syncwarp();
            int rk = 0;
            for (int i = 0; i < nsel; i++) {
                const i64 ci = (i64)(BS.Q[i] >> 9);
                bool gt = ci > mcnt;
                if (ci == mcnt && i != lane && sel != 0) {
                    const u64 ki = BS.Qkey[i];
                    const int32_t ai = (int32_t)((ki >> 32) & 0x7fffffff), bi = (int32_t)(ki & 0xffffffffu);
                    int r = tok_cmp_pre(M, ai, BS.Spa[i], ma, pa);
                    if (r == 0) r = tok_cmp_pre(M, bi, BS.Spb[i], mb, pb);
                    gt = r > 0;
                }
                if (gt) rk++;
            }
            if (sel != 0) pos = rk;
        }
        __syncwarp();
        if (lane < ML_SEL) {                                   // the selected entries in exact order (empty ones stay behind)
            BS.mem[pos].a = ma; BS.mem[pos].b = mb; BS.mem[pos].slot = mslot; BS.mem[pos].idx = midx; BS.mem[pos].cnt = sel != 0 ? mcnt : 0;
        }
        __syncwarp();
        if (lane < ML_SEL) { ma = BS.mem[lane].a; mb = BS.mem[lane].b; mslot = BS.mem[lane].slot; midx = BS.mem[lane].idx; mcnt = BS.mem[lane].cnt; }
        else mcnt = 0;
        const bool have = lane < nsel;                         // positions [0, nsel) are the entries (ranks are a permutation of them)
        SELT(3);
        // (callers batch only under a plain count threshold -- T2pa == 0 --, where the list holds EVERY pair with count >= T2)
        const bool elig = lane < batch_max && have && mcnt > g && mcnt >= T2 && mcnt >= T && mcnt >= Tmin;
#if ML_SEL_WHY
        if (lane == 0 && blockIdx.x == 0 && !elig) {          // why the best entry itself is not eligible: state[37..39] = no entry | tied beyond the selection | below T2, state[63] = below T (tuning build)
            const int why = !have ? 0 : (!(mcnt > g) ? 1 : (!(mcnt >= T2) ? 2 : (!(mcnt >= T) ? 3 : 4)));
            atomicAdd((u64*)&M.state[why < 3 ? 37 + why : 63], 1ULL);
        }
#endif
        int tj = 99;                                           // first earlier entry this one touches (or that has equal tokens)
        for (int i = nsel - 2; i >= 0; i--) {
            const int32_t ai = BS.mem[i].a, bi = BS.mem[i].b;
            if (i < lane && (ma == bi || mb == ai || ai == bi)) tj = i;
        }
        int k2 = __ffs(~__ballot_sync(0xffffffffu, elig && tj == 99)) - 1;      // members: eligible and clear of every earlier one
        // an entry left out with the count of the last member must not touch a member either (its count must stay
        // what it is, and it bounds the pairs the batch creates): give up members until that holds
        while (k2 > 1) {
            const i64 ck = BS.mem[k2 - 1].cnt;
            if (!__ballot_sync(0xffffffffu, lane >= k2 && have && mcnt == ck && tj < k2)) break;
            k2--;
        }
        if (lane == 0) BS.nb = k2;
        SELT(4);
    }
    __syncthreads();
    SELT(5);
    const int nb = BS.nb;
    if (threadIdx.x == 0) BS.nS = 0;                           // every thread read it before this barrier; the next call comes after another one
    return nb;
}


// ONE lane records one member of a batch and creates its token (members commit side by side in the lanes of one warp)
__device__ void commit_member(const MergeParams& M, LeaderCtx* lc, i64 m, int32_t a, int32_t b, int32_t c, bool is_new, i64 seg_base,
                              const MergedInfo& MI, i64 oc, bool publish) {
    M.merges[2 * m] = a; M.merges[2 * m + 1] = b; M.merge_new[m] = c;
    M.seg_start[m] = (int32_t)seg_base;
    M.merge_next[m] = is_new ? -1 : M.tok_first[c];
    if (!is_new) return;
    if (oc + MI.la + MI.lb > M.tok_bytes_cap || c + 1 >= M.max_tokens) {
        atomicOr((u64*)&M.state[MS_ERROR], (u64)ME_TOK_POOL_FULL); if (lc) lc->error = 1;
        return;
    }
    for (i64 k = 0; k < MI.la; k++) M.tok_bytes[oc + k] = M.tok_bytes[MI.oa + k];
    for (i64 k = 0; k < MI.lb; k++) M.tok_bytes[oc + MI.la + k] = M.tok_bytes[MI.ob + k];
    M.tok_off[c + 1] = oc + MI.la + MI.lb;
    M.tok_hash[c] = MI.H; M.tok_pow[c] = MI.P;
    M.tok_pre[c] = tok_prefix_concat(M.tok_pre[a], MI.la, M.tok_pre[b]);
    M.tok_first[c] = -1; M.tok_head[c].x = -1;
    // two members may have found the same free slot of the token set: claim it, move on when it is taken
    const u64 mask = (u64)M.tset_cap - 1, val = (MI.H & 0xffffffff00000000ULL) | (u64)(uint32_t)(c + 1);
    u64 slot = (u64)MI.tslot;
    while (atomicCAS(&M.tset[slot], 0ULL, val) != 0ULL) slot = (slot + 1) & mask;
    if (publish) { M.state[MS_NTOK] = c + 1; M.state[MS_POOL_USED] = oc + MI.la + MI.lb; }
}

// Stage C of a batch.  Lane l of every warp holds member l (pair, product token, candidate count, first item, log segment).
// The items of all members are laid end to end (every member starts at a multiple of the groups per warp, so the groups of a
// warp mostly work on the same member); a G-lane group claims its candidate word for the whole batch, finds out which
// members have a site in it (one pass over its symbols) and applies those, in order.  Software pipeline over the passes of a
// group as in leader_rewrite: the item of pass p+2 and the word header of pass p+1 are loaded while pass p is rewritten.
// Leader mode: lc = the leader's context (claims in its shared-memory set, counters in shared memory), group0 / ngroups span
// the rewriting warps of CTA 0.  Grid mode: lc = nullptr (claims by the per-word stamp, counters in global memory), the
// groups of ALL CTAs share the items.
template <int G>
__device__ __forceinline__ void rewrite_batch(const MergeParams& M, LeaderCtx* lc, const Ranges* RR, const u64* MK, int kk, int lane, int group0, int ngroups, int32_t stamp,
                                              int32_t ma, int32_t mb, int32_t mc, int32_t mslot, int mnew, int mtot, int mpst, int mseg,
                                              i64 T, i64 T2) {
    const int gl = lane & (G - 1), lead = lane & ~(G - 1);
    const int total_p = __shfl_sync(0xffffffffu, mpst + mtot, kk - 1);
    // Item `it` of the padded layout (all 32 lanes call this together).  The affected-word segment of a member is simply a
    // COPY of its candidate list -- a superset of the words it rewrites, which is all the index needs (candidates are ~97 %
    // hits) -- so the log needs no counter: hundreds of groups appending through one atomic serialised in the L2.
    auto fetch = [&](int it) -> i64 {
        // which member owns item `it`: lane l answers for member l, one ballot per group of the warp (not a loop over members)
        int mem = -1;
#pragma unroll
        for (int q = 0; q < 32 / G; q++) {
            const int itq = __shfl_sync(0xffffffffu, it, q * G);
            const uint32_t own = __ballot_sync(0xffffffffu, lane < kk && itq >= mpst && itq < mpst + mtot);
            if (lane / G == q) mem = own ? __ffs(own) - 1 : -1;
        }
        const int src = mem >= 0 ? mem : 0;
        const int loc = it - __shfl_sync(0xffffffffu, mpst, src), sb = __shfl_sync(0xffffffffu, mseg, src);
        const i64 e = mem >= 0 ? range_item(RR[mem], loc) : -1;
        if (mem >= 0 && gl == 0) M.alog_word[sb + loc] = e;
        return e;
    };
    auto claim = [&](int32_t w) -> int32_t {             // first claim of a word in this batch wins
        int take = 0;
        if (w >= 0 && gl == 0) take = lc ? (dedupe_claim(lc, w) ? 1 : 0) : (atomicExch(&M.wstamp[w], stamp) != stamp ? 1 : 0);
        take = __shfl_sync(0xffffffffu, take, lead);
        return take ? w : -1;
    };
    const int it0 = group0 + lane / G;
#if ML_RW_TRACE
    long long _t0 = clock64();
#endif
    const i64 e_cur = fetch(it0);
    i64 e_nx = fetch(it0 + ngroups);
    RWT(0, (int)e_cur + (int)e_nx);
    // the header (and the symbols: the entry carries their slot) is asked for BEFORE the claim returns: only the claimer
    // changes a word, so a read that loses the claim is merely dropped, and the two round trips overlap
    int32_t w_cur = POST_WORD(e_cur), w_nx = POST_WORD(e_nx);
    // ... and so are the word's first G symbols (almost every word fits): claim, header and symbols are ONE round trip.
    // (wsym has 8 slots of slack: a read past a short word's end returns its neighbour's symbols, masked by the length)
    uint32_t off_cur = 0; int n_cur = 0; i64 f_cur = 0; int32_t y0_cur = 0, y1_cur = 0;
    if (w_cur >= 0) {
        off_cur = (uint32_t)POST_OFF(e_cur);
        if (gl == 0) prefetch_l2(&M.wslot[off_cur]);
        n_cur = M.wlen[w_cur]; f_cur = M.wcnt[w_cur];
        y0_cur = M.wsym[off_cur + gl]; y1_cur = M.wsym[off_cur + gl + 1];
    }
    w_cur = claim(w_cur);
    RWT(1, w_cur + n_cur + (int)f_cur);
    for (int base = 0; base + group0 < total_p; base += ngroups) {      // warps without an item go straight to the barrier
        const i64 e_nx2 = fetch(base + it0 + 2 * ngroups);
        uint32_t off_nx = 0; int n_nx = 0; i64 f_nx = 0; int32_t y0_nx = 0, y1_nx = 0;
        if (w_nx >= 0) {
            off_nx = (uint32_t)POST_OFF(e_nx);
            if (gl == 0) prefetch_l2(&M.wslot[off_nx]);
            n_nx = M.wlen[w_nx]; f_nx = M.wcnt[w_nx];
            y0_nx = M.wsym[off_nx + gl]; y1_nx = M.wsym[off_nx + gl + 1];
        }
        w_nx = claim(w_nx);
        // ---- the current word: which members have a site in it
        const int32_t w = w_cur;
        const i64 off = (i64)off_cur;
        int n = w >= 0 ? n_cur : 0;
        uint32_t mask = 0;
        {
            const int32_t* s = M.wsym + off;
            for (int j0 = 0; __any_sync(0xffffffffu, j0 + 1 < n); j0 += G) {
                const int j = j0 + gl;
                u64 key = 0;
                if (j + 1 < n) key = j0 == 0 ? PAIR_KEY(y0_cur, y1_cur) : PAIR_KEY(s[j], s[j + 1]);
#pragma unroll 4
                for (int i = 0; i < kk; i++) mask |= key == MK[i] ? 1u << i : 0u;        // broadcast loads, independent of each other
            }
#pragma unroll
            for (int o = 1; o < G; o <<= 1) mask |= __shfl_xor_sync(0xffffffffu, mask, o);
        }
        RWT(2, mask + w_nx + n_nx);
        // members with a site in any word of this warp, in order (usually one)
        for (uint32_t todo = __reduce_or_sync(0xffffffffu, w >= 0 ? mask : 0u); todo; todo &= todo - 1) {
            const int i = __ffs(todo) - 1;
            const bool act = w >= 0 && ((mask >> i) & 1u);
            const int32_t a = __shfl_sync(0xffffffffu, ma, i), b = __shfl_sync(0xffffffffu, mb, i), c = __shfl_sync(0xffffffffu, mc, i);
            const int32_t sl = __shfl_sync(0xffffffffu, mslot, i);
            const int isn = __shfl_sync(0xffffffffu, mnew, i);
            if (a != b) {
                const int nn = rewrite_words_g<G>(M, act ? w : -1, off, n, f_cur, a, b, c, T, T2, lc, isn != 0, nullptr, -1);
                if (act) n = nn;
            } else if (act && gl == 0) rewrite_word_thread(M, w, a, b, c, T, T2, lc, isn != 0, sl, nullptr, -1);
        }
        RWT(3, n);
        w_cur = w_nx; off_cur = off_nx; n_cur = n_nx; f_cur = f_nx; y0_cur = y0_nx; y1_cur = y1_nx; e_nx = e_nx2; w_nx = POST_WORD(e_nx2);
    }
}

// Leader loop.  Per iteration (every stage is bounded by dependent L2 / DRAM round trips, not by bandwidth):
//   A  the ML_SEL best entries of the top list (counts mirrored in shared memory) -> the batch; ties at the top: byte-wise argmax
//   B  candidate ranges (one thread per member) and merged-token lookups (another thread per member), side by side
//   C  the last warp records the merges / creates the tokens while every other 8-lane group takes ONE candidate
//      item, claims its word in a shared-memory set (no global stamp), loads the word and rewrites it
//   D  the members' affected-log segments are closed (plain stores; all counters are in shared memory), thresholds of the new pairs
__device__ void leader_loop(const MergeParams& M, LeaderCtx& C, BatchCtx& BC, Best* sh_best, i64 T, i64 Tmin, const i64 T2, const u64 T2pa) {
    __builtin_assume(__isShared(&BC));          // a reference parameter hides the address space: without this every access is a generic load
    __builtin_assume(__isShared(&C));
    __shared__ i64 sh_wmax[ML_THREADS / 32];    // block-wide maxima of the tie path
#if ML_TIMING
    __shared__ long long sh_tacc[8];
    if (threadIdx.x < 8) sh_tacc[threadIdx.x] = 0;
#endif
    __shared__ int sh_ncand;
    __shared__ Best sh_cand[32];
    for (int i = threadIdx.x; i < 2 * ML_TOP_N; i += blockDim.x) C.LM.mkey[i] = 0;
    for (int i = threadIdx.x; i < ML_DEDUPE_N; i += blockDim.x) C.dedupe[i] = 0;
    i64 pool_end = M.tok_off[(int32_t)M.state[MS_NTOK]];
    i64 m = M.state[MS_NMERGES];
    int32_t n_tok = (int32_t)M.state[MS_NTOK];
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, lane = threadIdx.x & 31;
    int batch_max = (M.batch_max & 255) > 0 ? (int)(M.batch_max & 255) : ML_BATCH_MAX;
    if (batch_max > 16) batch_max = 16;  // its batches are bounded by ML_LEADER_BATCH_ITEMS anyway; a head of 32 keeps the selection short
    const bool grid_batches_on = ((M.batch_max >> 8) & 255 ? (M.batch_max >> 8) & 255 : M.batch_max & 255) != 1 && !((M.batch_max >> 17) & 1);   // bit 17: keep big batches here (tuning)
    if (T2pa != 0) batch_max = 1;        // tie regime: the list is not complete at the boundary count (and pair_add reads prefixes of tokens being created)
#if ML_TIMING
    long long t_arg = 0, t_rng = 0, t_rw = 0, t_close = 0;
#endif
    int s_act = 0, s_items = 0, n_done = 0, n_iter = 0, n_batched = 0;
    int reason = LR_OTHER;
    int cached = 0;                      // entries of the top list whose count is mirrored in shared memory
    i64 theta = 0;                       // head threshold (0: not chosen yet)
    bool theta_sticky = false;
    const i64 npairs0 = __ldcg(&M.state[MS_NPAIRS]);
    const i64 last_rebuild_m = __ldcg(&M.state[MS_LAST_REBUILD_M]);
    if (threadIdx.x == 0) {
        C.act_n = (int)__ldcg(&M.state[MS_ACT_N]); C.alog_n = (int)__ldcg(&M.state[MS_ALOG_N]);
        C.error = (int)__ldcg(&M.state[MS_ERROR]);
        C.top_n = (int)__ldcg(&M.state[MS_TOP_N]); C.top_ovf = (int)__ldcg(&M.state[MS_TOP_OVF]);
        C.npairs_new = 0; C.nnew = 0; C.t2pa = T2pa;
        sh_ncand = 0; BC.sel.nS = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) *(volatile i64*)&M.state[MS_TOP_N_LIVE] = C.top_n;
    {
        const int tn0 = C.top_n < ML_TOP_N ? C.top_n : ML_TOP_N;
        if ((int)threadIdx.x < tn0) { C.tslot[threadIdx.x] = M.top_slot[threadIdx.x]; C.tkey[threadIdx.x] = M.top_key[threadIdx.x]; C.LM.pending[threadIdx.x] = 1; }
    }
    __syncthreads();
    for (int iter = 0; iter < ML_LEADER_BATCH && m < M.num_merges; iter++) {
        ML_CLOCK(c0);
        ML_T0(qa);
        ML_TR(0);
        const int alog_n = C.alog_n;
        if (C.error) break;
        if (C.top_ovf || C.top_n > ML_TOP_N) { reason = LR_TOP; break; }
        if (M.rebuild_every > 0 && m - last_rebuild_m >= M.rebuild_every) { reason = LR_REBUILD; break; }
        if ((npairs0 + C.npairs_new + C.nnew) * 4 > M.pcap * 3) { if (threadIdx.x == 0) atomicOr((u64*)&M.state[MS_ERROR], (u64)ME_PAIR_TABLE_FULL); break; }
        // ---- A: the best pairs of the top list
        const int tn = C.top_n;
        Best mine{0, -1, 0, 0, 0};
        if ((int)threadIdx.x < tn) {
            const int32_t sl = C.tslot[threadIdx.x];
            const u64 k = C.tkey[threadIdx.x];
            if ((int)threadIdx.x >= cached && C.LM.pending[threadIdx.x]) {   // entry without a mirrored count yet
                mirror_set(&C.LM, threadIdx.x, __ldcg(&M.pcnt[sl]));
                mirror_insert(&C.LM, sl, threadIdx.x);
                const int32_t ka = (int32_t)((k >> 32) & 0x7fffffff), kb = (int32_t)(k & 0xffffffffu);
                C.tpa[threadIdx.x] = M.tok_pre[ka]; C.tpb[threadIdx.x] = M.tok_pre[kb];
                C.tha[threadIdx.x] = M.tok_hash[ka]; C.thb[threadIdx.x] = M.tok_hash[kb]; C.tpwb[threadIdx.x] = M.tok_pow[kb];
                const uint32_t p0 = M.ioff[sl], p1 = M.ioff[sl + 1];
                C.tp0[threadIdx.x] = p0; C.tplen[threadIdx.x] = p1 - p0;
            }
            const i64 cnt = mirror_get(&C.LM, threadIdx.x);
            if (cnt > 0) mine = Best{cnt, sl, (int32_t)((k >> 32) & 0x7fffffff), (int32_t)(k & 0xffffffffu), (int32_t)threadIdx.x};
        }
        cached = tn;
        ML_TACC(0, qa);                       // (timing builds) stage A in three parts: entries | batch selection | one-merge argmax
        // member registers: lane l (< ML_BATCH_MAX) of EVERY warp describes member l of the batch
        int32_t ma = 0, mb = 0, mslot = -1, midx = -1; i64 mcnt = 0;
        int nb = 0;
        Best best{0, -1, 0, 0, 0};
        if (batch_max > 1) {
            nb = select_batch(M, BC.sel, mine, tn, batch_max, T, Tmin, T2, C.tpa, C.tpb, sh_wmax, theta, theta_sticky, 32);
            if (lane < ML_BATCH_MAX) { ma = BC.sel.mem[lane].a; mb = BC.sel.mem[lane].b; mslot = BC.sel.mem[lane].slot; midx = BC.sel.mem[lane].idx; mcnt = BC.sel.mem[lane].cnt; }
            if (nb > 0) best = Best{__shfl_sync(0xffffffffu, mcnt, 0), __shfl_sync(0xffffffffu, mslot, 0), __shfl_sync(0xffffffffu, ma, 0),
                                    __shfl_sync(0xffffffffu, mb, 0), __shfl_sync(0xffffffffu, midx, 0)};
        }
        ML_TACC(1, qa);
        if (nb == 0) {
            // ---- one merge: the maximum of the whole list, (left bytes, right bytes) among equal counts
            best = leader_argmax_one(M, C, mine, sh_wmax, &sh_ncand, sh_cand, sh_best);
            if (best.slot >= 0) { nb = 1; ma = best.a; mb = best.b; mslot = best.slot; midx = best.pad; mcnt = best.cnt; }
        }
        ML_TACC(2, qa);
        if (best.slot < 0 || best.cnt < T2 || (best.cnt == T2 && best.pad >= 0 && C.tpa[best.pad] < T2pa)) { reason = LR_TOP; break; }
        if (best.cnt < T || best.cnt < Tmin) break;                        // threshold step / termination: grid mode
        if (m + nb > M.num_merges) nb = (int)(M.num_merges - m);
        ML_CLOCK(c1);
        ML_TR(1);
        // ---- B: candidate ranges + merged tokens, one thread each per member (ranges: warp i, lookups: warp 31 - i; lane i holds member i)
        if (warp < nb && lane == warp) build_ranges(M, mslot, ma, mb, &BC.R[warp], midx >= 0, midx >= 0 ? C.tp0[midx] : 0u, midx >= 0 ? C.tplen[midx] : 0u);
        if (nwarps - 1 - warp < nb && lane == nwarps - 1 - warp)
            BC.c[lane] = lookup_merged_leader(M, C, midx, ma, mb, n_tok, &BC.MI[lane]);
        if (warp == nwarps / 2 && lane < nb) BC.mkey[lane] = PAIR_KEY(ma, mb);
        for (int i = threadIdx.x; i < ML_DEDUPE_N; i += blockDim.x) C.dedupe[i] = 0;
        __syncthreads();
        if (threadIdx.x == 0) { sh_ncand = 0; C.npairs_new += C.nnew; C.nnew = 0; }   // everybody has read them; next use is after stage C's barrier
        // how many members fit: candidate lists, log space, distinct NEW product tokens (every warp computes the same)
        int mtot = 0, mnew = 0, mlen = 0; int32_t mc = -1; u64 mH = 0; bool mbad = true;
        if (lane < nb) {
            mtot = (int)(BC.R[lane].total < 0x40000000 ? BC.R[lane].total : 0x40000000);
            mbad = BC.R[lane].n < 0;
            mc = BC.c[lane]; mnew = mc == n_tok ? 1 : 0; mH = BC.MI[lane].H; mlen = (int)(BC.MI[lane].la + BC.MI[lane].lb);
            if (mnew) mc = n_tok + lane;                 // every earlier member makes a new token (or the batch ends there)
        }
        int cum = mtot, lcum = mlen;                                 // inclusive scans over the members
#pragma unroll
        for (int o = 1; o < ML_BATCH_MAX; o <<= 1) {
            const int t1 = __shfl_up_sync(0xffffffffu, cum, o), t3 = __shfl_up_sync(0xffffffffu, lcum, o);
            if (lane >= o) { cum += t1; lcum += t3; }
        }
        if (cum > ML_LEADER_ITEMS_MAX || (i64)alog_n + cum > M.alog_cap) mbad = true;
        for (int i = 0; i < nb - 1; i++) {                           // (nothing to do for the common single merge)
            const u64 Hi = __shfl_sync(0xffffffffu, mH, i);
            const int newi = __shfl_sync(0xffffffffu, mnew, i);
            if (i < lane && (Hi == mH || !newi)) mbad = true;      // same merged bytes twice / an earlier member reuses an existing token
        }
        const int kk = __ffs(__ballot_sync(0xffffffffu, mbad)) - 1;     // lanes >= nb are bad: kk <= nb
        if (kk == 0) break;                                              // the best pair alone needs the grid (or an index rebuild)
        // A batch with many candidate words is better off in grid mode: one pass over 148 SMs instead of several passes here
        // (the members are the same there; the main loop stays in grid mode with exponential back-off while that goes on).
        if (nb >= 2 && grid_batches_on && __shfl_sync(0xffffffffu, cum, nb - 1) > ML_LEADER_BATCH_ITEMS) break;
        ML_CLOCK(c2);
        ML_TR(2);
        int n_new_batch = 0, len_batch = 0, items_batch = 0;
        if (kk == 1) {
            // ---- one merge (the path of every heavy, tied or interacting pair)
            const Ranges& R = BC.R[0];
            const MergedInfo& MI = BC.MI[0];
            const int32_t a = best.a, b = best.b, c = BC.c[0];
            const bool is_new = c == n_tok;
            // ---- C: commit (last warp) || claim + rewrite (one 8-lane group per candidate item)
            if (warp == nwarps - 1) commit_merge_leader(M, m, a, b, c, is_new, alog_n, MI, pool_end);
            else {
                // 8-lane groups (124 candidates per pass) while that is one pass, 4-lane groups (248 per pass) beyond
                if (a != b && R.total > (nwarps - 1) * 8) leader_rewrite<2>(M, C, R, warp, lane, nwarps, a, b, c, T, T2, is_new, best.slot);
                else if (a != b && R.total > (nwarps - 1) * 4) leader_rewrite<4>(M, C, R, warp, lane, nwarps, a, b, c, T, T2, is_new, best.slot);
                else leader_rewrite<8>(M, C, R, warp, lane, nwarps, a, b, c, T, T2, is_new, best.slot);
            }
            n_new_batch = is_new ? 1 : 0; len_batch = is_new ? (int)(MI.la + MI.lb) : 0; items_batch = (int)R.total;
            ML_T0(qb);
            ML_TR(7);
            __syncthreads();
            ML_TACC(4, qb);
            ML_CLOCK(c3);
            ML_TR(8);
            // ---- D: close the merge; thresholds of the pairs it created
            if (threadIdx.x == 0) {
                const int32_t prev = is_new ? -1 : M.tok_first[c];        // == merge_next[m] written by the commit warp
                M.seg_end[m] = C.alog_n; M.tok_first[c] = (int32_t)m;
                M.tok_head[c] = make_int4((int32_t)m, alog_n, C.alog_n, prev);
                *(volatile i64*)&M.state[MS_TOP_N_LIVE] = C.top_n;        // for the prefetch helpers
            }
            if (threadIdx.x == 32) { M.pcnt[best.slot] = 0; const int bi = mirror_find(&C.LM, best.slot); if (bi >= 0) mirror_set(&C.LM, bi, 0); }
            if (is_new) leader_new_pairs(M, &C, T, T2);
            __syncthreads();
#if ML_TIMING
            { ML_CLOCK(c4); t_arg += c1 - c0; t_rng += c2 - c1; t_rw += c3 - c2; t_close += c4 - c3; }
#endif
        } else {
            // ---- a batch of kk merges
            // lanes per candidate word: 8 while that is one pass (124 items), then 4, then 2 (496 items a pass)
            const int items_all = __shfl_sync(0xffffffffu, cum, kk - 1);
            const int G = items_all > (nwarps - 1) * 8 ? 2 : (items_all > (nwarps - 1) * 4 ? 4 : 8), padm = 32 / G - 1;
            int pcum = lane < kk ? (mtot + padm) & ~padm : 0;                                 // every member starts at a multiple of the groups per warp
#pragma unroll
            for (int o = 1; o < ML_BATCH_MAX; o <<= 1) { const int t2 = __shfl_up_sync(0xffffffffu, pcum, o); if (lane >= o) pcum += t2; }
            const int mpst = pcum - (lane < kk ? (mtot + padm) & ~padm : 0), mseg = alog_n + cum - mtot;     // first item (padded layout), log segment
            const i64 moc = pool_end + lcum - mlen;                                          // where the member's token bytes go
            if (threadIdx.x == kk - 1) C.alog_n = alog_n + cum;                              // the members' segments are reserved in full
            // ---- C: commits (last warp, one lane per member) || claim + rewrite
            const int last_new = __shfl_sync(0xffffffffu, mnew, kk - 1) ? kk - 1 : kk - 2;   // every member but the last makes a new token
            if (warp == nwarps - 1) {
                if (lane < kk) commit_member(M, &C, m + lane, ma, mb, mc, mnew != 0, (i64)mseg, BC.MI[lane], moc, lane == last_new);
            } else if (G == 8) rewrite_batch<8>(M, &C, BC.R, BC.mkey, kk, lane, warp * 4, (nwarps - 1) * 4, 0, ma, mb, mc, mslot, mnew, mtot, mpst, mseg, T, T2);
            else if (G == 4) rewrite_batch<4>(M, &C, BC.R, BC.mkey, kk, lane, warp * 8, (nwarps - 1) * 8, 0, ma, mb, mc, mslot, mnew, mtot, mpst, mseg, T, T2);
            else rewrite_batch<2>(M, &C, BC.R, BC.mkey, kk, lane, warp * 16, (nwarps - 1) * 16, 0, ma, mb, mc, mslot, mnew, mtot, mpst, mseg, T, T2);
            n_new_batch = __popc(__ballot_sync(0xffffffffu, lane < kk && mnew));
            len_batch = __shfl_sync(0xffffffffu, lcum, kk - 1) - (__shfl_sync(0xffffffffu, mnew, kk - 1) ? 0 : __shfl_sync(0xffffffffu, mlen, kk - 1));
            items_batch = __shfl_sync(0xffffffffu, cum, kk - 1);
            __syncthreads();
            ML_CLOCK(c3);
            // ---- D: close the members; thresholds of the pairs the batch created
            if (warp == 0 && lane < kk) {
                const int32_t prev = mnew ? -1 : M.tok_first[mc];
                const int e = mseg + mtot;
                M.seg_end[m + lane] = e; M.tok_first[mc] = (int32_t)(m + lane);
                M.tok_head[mc] = make_int4((int32_t)(m + lane), mseg, e, prev);
                M.pcnt[mslot] = 0; mirror_set(&C.LM, midx, 0);
                if (lane == 0) *(volatile i64*)&M.state[MS_TOP_N_LIVE] = C.top_n;
            }
            leader_new_pairs(M, &C, T, T2);
            __syncthreads();
#if ML_TIMING
            { ML_CLOCK(c4); t_arg += c1 - c0; t_rng += c2 - c1; t_rw += c3 - c2; t_close += c4 - c3; }
#endif
            n_batched += kk;
        }
        ML_TR(9);
        s_act += C.alog_n - alog_n; s_items += items_batch; n_done += kk; n_iter++;
        m += kk; n_tok += n_new_batch; pool_end += len_batch;
    }
    __syncthreads();
    // the mirror must agree with the table (cheap self-check, once per leader session)
    if ((int)threadIdx.x < cached && mirror_get(&C.LM, threadIdx.x) != __ldcg(&M.pcnt[C.tslot[threadIdx.x]]))
        atomicOr((u64*)&M.state[MS_ERROR], (u64)ME_INTERNAL);
    if (threadIdx.x == 0) {
        M.state[MS_NMERGES] = m;
        M.state[MS_ACT_N] = C.act_n; M.state[MS_ALOG_N] = C.alog_n;
        M.state[MS_TOP_N] = C.top_n; M.state[MS_TOP_OVF] = C.top_ovf;
        if (C.npairs_new + C.nnew) atomicAdd((u64*)&M.state[MS_NPAIRS], (u64)(C.npairs_new + C.nnew));
#if ML_TIMING
        M.state[20] += t_arg; M.state[21] += t_rng; M.state[23] += t_rw; M.state[24] += t_close;
        M.state[37] += sh_tacc[0]; M.state[38] += sh_tacc[1]; M.state[39] += sh_tacc[2];
#endif
        M.state[25] += s_act; M.state[26] += s_items;
        M.state[MS_LEADER_MERGES] += n_done;
        M.state[MS_LEADER_ITERS] += n_iter; M.state[MS_LEADER_BATCHED] += n_batched;
        M.state[MS_LEADER_REASON] = reason;
        if (reason == LR_TOP) M.state[MS_T2] = 0;
    }
}

// grid-wide rebuild of the top list.  On return state[MS_T2] holds the new threshold (> 0), or -1 when
// more than ML_TOP_N pairs tie near the maximum (the caller then scans the whole active set for one
// merge).  Lowers T (and rebuilds the active set) while no pair reaches it; returns false when no pair
// with count >= Tmin is left.  Must be entered by all CTAs right after a grid barrier.
__device__ bool grid_top_rebuild(const MergeParams& M, i64& T, i64 Tmin, Best* sh_best, int* sh_hist) {
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gstride = (i64)gridDim.x * blockDim.x;
    {
        i64 n_old = M.state[MS_TOP_N]; if (n_old > ML_TOP_N) n_old = ML_TOP_N;
        for (i64 i = gtid; i < n_old; i += gstride) { int32_t sl = M.top_slot[i]; atomicAnd(&M.intop[sl >> 5], ~(1u << (sl & 31))); }
        for (i64 i = gtid; i < 1024; i += gstride) M.hist[i] = 0;
    }
    grid_barrier(M);
    if (gtid == 0) { M.state[MS_TOP_N] = 0; M.state[MS_TOP_OVF] = 0; M.state[MS_TOP_REBUILDS]++; }
    Best am;
    for (;;) {
        am = grid_argmax(M, sh_best);              // contains a grid barrier
        if (am.slot >= 0 && am.cnt >= T) break;
        if (T <= Tmin) return false;
        T = T / 2; if (T < Tmin) T = Tmin;
        grid_barrier(M);
        rebuild_active(M, T);
    }
    const i64 hi = am.cnt, span = hi - T, act_n = M.state[MS_ACT_N];
    for (i64 i = gtid; i < act_n; i += gstride) {
        i64 c = __ldcg(&M.pcnt[M.act[i]]);
        if (c < T) continue;
        int bin = span > 0 ? (int)(((c - T) * 1023) / span) : 1023;
        atomicAdd(&M.hist[bin], 1);
    }
    grid_barrier(M);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh_hist[i] = M.hist[i];
    __syncthreads();
    i64 T2;
    u64 PA = 0;
    {
        int acc = 0, b = 1023;
        while (b >= 0 && acc + sh_hist[b] <= (ML_TOP_N * 3) / 4) { acc += sh_hist[b]; b--; }
        T2 = b < 0 ? T : (b >= 1023 ? hi : T + ((i64)(b + 1) * span + 1022) / 1023);
        if (T2 < T) T2 = T;
        // Every bin is ONE count when the span is small (the tie regime).  The bin that did not fit is then cut by the left
        // token's prefix: radix select (8 bits a level, most significant first) of about the `want` largest prefixes among
        // the pairs with exactly that count.
        if (b >= 0 && span <= 1022 && sh_hist[b] > 0) {
            const i64 cb = span > 0 ? T + ((i64)b * span + 1022) / 1023 : hi;          // the count of bin b
            const bool exact = span == 0 || (int)(((cb - T) * 1023) / span) == b;
            int want = (ML_TOP_N * 3) / 4 - acc, room = ML_TOP_N - 32 - acc;
            if (exact && want > 0) {
                u64 prefix = 0;
                bool done = false;
                for (int level = 0; level < 8 && !done; level++) {
                    const int shift = 56 - 8 * level;
                    for (i64 i = gtid; i < 256; i += gstride) M.hist[i] = 0;
                    grid_barrier(M);
                    for (i64 i = gtid; i < act_n; i += gstride) {
                        const int32_t sl = M.act[i];
                        if (__ldcg(&M.pcnt[sl]) != cb) continue;
                        const u64 pre = M.tok_pre[(int32_t)((__ldcg(&M.pkey[sl]) >> 32) & 0x7fffffff)];
                        if (level > 0 && (pre >> (shift + 8)) != (prefix >> (shift + 8))) continue;
                        atomicAdd(&M.hist[(int)((pre >> shift) & 255)], 1);
                    }
                    grid_barrier(M);
                    if (threadIdx.x < 256) sh_hist[threadIdx.x] = __ldcg(&M.hist[threadIdx.x]);
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        int cum = 0, chosen = -1;
                        for (int bin = 255; bin >= 0; bin--) {
                            const int h = sh_hist[bin];
                            if (cum + h > want) { chosen = bin; break; }
                            cum += h;
                        }
                        sh_hist[256] = chosen; sh_hist[257] = cum; sh_hist[258] = chosen >= 0 ? sh_hist[chosen] : 0;
                    }
                    __syncthreads();
                    const int chosen = sh_hist[256], cum = sh_hist[257], hc = sh_hist[258];
                    grid_barrier(M);                                  // everybody has read the histogram before it is zeroed again
                    if (chosen < 0) { done = true; break; }             // everything under this prefix fits
                    prefix |= (u64)chosen << shift;
                    want -= cum; room -= cum;
                    if (hc <= room || level == 7) { done = true; break; }   // take the whole bin (level 7: equal prefixes cannot be split)
                }
                T2 = cb; PA = prefix;
            }
        }
    }
    for (i64 i = gtid; i < act_n; i += gstride) {
        int32_t sl = M.act[i];
        const i64 c = __ldcg(&M.pcnt[sl]);
        if (c < T2) continue;
        const u64 key = __ldcg(&M.pkey[sl]);
        if (c == T2 && PA != 0 && M.tok_pre[(int32_t)((key >> 32) & 0x7fffffff)] < PA) continue;
        i64 idx = (i64)atomicAdd((u64*)&M.state[MS_TOP_N], 1ULL);
        if (idx < ML_TOP_N) { M.top_slot[idx] = sl; M.top_key[idx] = key; atomicOr(&M.intop[sl >> 5], 1u << (sl & 31)); }
        else M.state[MS_TOP_OVF] = 1;
    }
    grid_barrier(M);
    const bool ovf = M.state[MS_TOP_OVF] != 0;
    grid_barrier(M);                                         // everyone has seen the overflow flag
    if (ovf) {
        i64 n_old = M.state[MS_TOP_N]; if (n_old > ML_TOP_N) n_old = ML_TOP_N;
        for (i64 i = gtid; i < n_old; i += gstride) { int32_t sl = M.top_slot[i]; atomicAnd(&M.intop[sl >> 5], ~(1u << (sl & 31))); }
        grid_barrier(M);
        if (gtid == 0) { M.state[MS_TOP_N] = 0; M.state[MS_TOP_OVF] = 0; M.state[MS_T2] = -1; M.state[MS_T2_PA] = 0; M.state[MS_TIE_CNT] = hi; M.state[MS_TIE_LEFT] = 64; }
    } else if (gtid == 0) { M.state[MS_T2] = T2; M.state[MS_T2_PA] = (i64)PA; }
    grid_barrier(M);
    return true;
}

// ---- prefetch helpers (leader mode) ---------------------------------------------------------------------------------
// While CTA 0 runs merges alone, the word arrays of a large corpus (GBs) miss the L2, so every dependent step of a merge
// (postings -> word header -> symbols) is a DRAM round trip.  The pair that will be merged NEXT is almost always among the
// few largest of the top list, which lives in global memory: a helper CTA (otherwise idle at the grid barrier) keeps
// finding the ML_SEL largest entries, walks their candidate lists exactly as the leader will, and pulls the lines the
// leader is going to touch into the L2 (prefetch.global.L2).  Helpers only read (through the L2: ld.cg) and prefetch;
// whatever they see -- stale list entries, half-updated words -- can only make a prefetch useless, never change a result.
#ifndef ML_HELPERS
#define ML_HELPERS 2
#endif
#ifndef ML_HELPER_MIN_SYMS
#define ML_HELPER_MIN_SYMS (8 << 20)      // word arrays below ~100 MB stay in the L2 anyway (126 MB): nothing to prefetch
#endif
__device__ void helper_loop(const MergeParams& M, i64 gen, int hidx, u64* sh_keys, Ranges* R) {
    __shared__ int sh_stop;
    for (int round = 0;; round++) {
        if (threadIdx.x == 0) sh_stop = *(volatile i64*)&M.state[MS_LEADER_GEN] != gen;
        __syncthreads();
        if (sh_stop) break;
        i64 tn = __ldcg(&M.state[MS_TOP_N_LIVE]);
        if (tn > ML_TOP_N) tn = ML_TOP_N;
        u64 key = 0;
        if ((i64)threadIdx.x < tn) {
            const i64 cnt = __ldcg(&M.pcnt[__ldcg(&M.top_slot[threadIdx.x])]);
            if (cnt > 0) key = ((u64)cnt << 9) | (u64)(511 - (int)threadIdx.x);
        }
        const u64 mykey = select_top(key, sh_keys);      // one block barrier inside; lane r holds the r-th largest key
        for (int r = hidx; r < ML_HSEL; r += ML_HELPERS) {
            __syncthreads();
            const u64 kr = __shfl_sync(0xffffffffu, mykey, r);
            if (kr == 0) continue;                       // block-uniform
            const int idx = 511 - (int)(kr & 511);
            if (threadIdx.x == 0) {
                const int32_t slot = __ldcg(&M.top_slot[idx]);
                const u64 k = __ldcg(&M.top_key[idx]);
                build_ranges(M, slot, (int32_t)((k >> 32) & 0x7fffffff), (int32_t)(k & 0xffffffffu), R);
            }
            __syncthreads();
            if (R->n < 0) continue;
            const i64 total = R->total < 2 * ML_LEADER_ITEMS_MAX ? R->total : 2 * ML_LEADER_ITEMS_MAX;
            for (i64 it = threadIdx.x; it < total; it += blockDim.x) {
                i64 j = it; int32_t w = -1;
                for (int q = 0; q < R->n; q++) { if (j < R->len[q]) { w = POST_WORD(__ldcg(&R->base[q][j])); break; } j -= R->len[q]; }
                if (w < 0 || w >= M.n_words) continue;
                prefetch_l2(&M.wcnt[w]);
                const i64 off = __ldcg(&M.woff[w]);
                const int n = __ldcg(&M.wlen[w]);
                if (off < 0 || off >= M.n_syms) continue;
                prefetch_l2(&M.wsym[off]); prefetch_l2(&M.wslot[off]);
                if (n > 8) { prefetch_l2(&M.wsym[off + (n < 64 ? n : 64) - 1]); prefetch_l2(&M.wslot[off + (n < 64 ? n : 64) - 1]); }
            }
        }
        __syncthreads();
        __nanosleep(500);
    }
}

extern __shared__ __align__(16) unsigned char ml_dyn_smem[];      // LeaderCtx (used by CTA 0 in leader mode)
#define ML_DYN_SMEM_BYTES ((int)sizeof(LeaderCtx))

__global__ void __launch_bounds__(ML_THREADS) k_merge_loop(MergeParams M) {
    __shared__ Best sh_best[ML_THREADS / 32];
    __shared__ i64 sh_scan[1 + ML_THREADS / 32];
    __shared__ i64 sh_cnt[ML_THREADS / 32];
    __shared__ __align__(16) int sh_hist[1024];
    __shared__ int32_t sh_c;
    __shared__ Ranges R;
    __shared__ BatchCtx GB;                  // grid-mode batches (every CTA builds the same one)
    __shared__ long long sh_phase[24];       // phase clocks of CTA 0: state[40 .. 63]
    if (threadIdx.x < 24) sh_phase[threadIdx.x] = 0;
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gstride = (i64)gridDim.x * blockDim.x;
    i64 g_theta = 0; bool g_sticky = false;   // head threshold of the grid-mode batch selection: the same in every CTA
    __shared__ int32_t g_tslot[ML_TOP_N];     // this CTA's copy of the top list (grid-mode batches), valid below g_cached
    __shared__ u64 g_tkey[ML_TOP_N];
    __shared__ u64 g_tpa[ML_TOP_N], g_tpb[ML_TOP_N];      // 8-byte prefixes of the entries' tokens
    int g_cached = 0;
    if (threadIdx.x == 0) GB.sel.nS = 0;

    long long pclk = clock64();
    const long long pclk0 = pclk;
    // ---- K4: pair histogram (trainer.py:228-235)
    for (i64 i = gtid; i + 1 < M.n_syms; i += gstride) {
        int32_t w = M.sym_word[i];
        i64 j = i - M.woff[w];
        if (j + 1 < M.wlen[w]) {
            i64 s = pair_upsert(M, PAIR_KEY(M.wsym[i], M.wsym[i + 1]));
            if (s >= 0) { atomicAdd((u64*)&M.pcnt[s], (u64)M.wcnt[w]); M.wslot[i] = (int32_t)s; }
        }
    }
    if (gtid == 0) { unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); M.state[MS_LEADER_SMID] = smid; }
    for (i64 t = gtid; t < M.max_tokens; t += gstride) { M.tok_first[t] = -1; M.tok_head[t] = make_int4(-1, 0, 0, -1); }
    for (i64 t = gtid; t < M.state[MS_NTOK]; t += gstride) M.tok_pre[t] = tok_prefix_of_bytes(M.tok_bytes + M.tok_off[t], M.tok_off[t + 1] - M.tok_off[t]);
    grid_barrier(M);
    {
        i64 mx = 0;
        for (i64 s = gtid; s < M.pcap; s += gstride) { i64 c = M.pcnt[s]; if (c > mx) mx = c; }
        for (int o = 16; o > 0; o >>= 1) { i64 t = __shfl_xor_sync(0xffffffffu, mx, o); if (t > mx) mx = t; }
        if ((threadIdx.x & 31) == 0 && mx > 0) atomicMax((i64*)&M.state[MS_MAXCNT], mx);
    }
    ML_PHASE(MS_CLK_HIST, pclk);
    rebuild_index(M, sh_scan, 0);      // starts and ends with grid-wide syncs
    const i64 Tmin = M.min_freq > 1 ? M.min_freq : 1;
    i64 T = M.state[MS_MAXCNT] / 2; if (T < Tmin) T = Tmin;
    rebuild_active(M, T);
    ML_PHASE(MS_CLK_INDEX0, pclk);

    // prefetch helpers of the leader mode (result-neutral): which CTAs, if any
    int helper_idx = -1;
    if (M.helper_mode != 0 && M.n_syms > M.helper_min_syms && blockIdx.x != 0) {
        if (M.helper_mode == 1) { if ((int)blockIdx.x <= ML_HELPERS) helper_idx = (int)blockIdx.x - 1; }
        else {
            unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            const unsigned lead = (unsigned)M.state[MS_LEADER_SMID];
            if (smid == (lead ^ 1u)) helper_idx = 0; else if (smid == (lead ^ 2u)) helper_idx = 1;
        }
    }

    int skip = 0, backoff = 1;
    for (;;) {
        // ---- every CTA reads the shared state right after a grid barrier
        i64 m = M.state[MS_NMERGES];
        int32_t n_tok = (int32_t)M.state[MS_NTOK];
        i64 alog_n = M.state[MS_ALOG_N];
        const i64 act_n = M.state[MS_ACT_N];
        const i64 gen = M.state[MS_LEADER_GEN];
        i64 T2 = M.state[MS_T2];
        const u64 T2pa = (u64)M.state[MS_T2_PA];
        const i64 tie_cnt = M.state[MS_TIE_CNT], tie_left = M.state[MS_TIE_LEFT];
        const i64 top_n = M.state[MS_TOP_N];
        const bool top_ovf = M.state[MS_TOP_OVF] != 0;
        if (m >= M.num_merges || M.state[MS_ERROR] || M.state[MS_DONE]) break;
        if (M.state[MS_NPAIRS] * 4 > M.pcap * 3) { if (gtid == 0) atomicOr((u64*)&M.state[MS_ERROR], (u64)ME_PAIR_TABLE_FULL); break; }

        // ---- (re)build the top list when it cannot prove the maximum any more
        if (T2 == 0 || (T2 > 0 && top_ovf)) {
            grid_barrier(M);                                    // everyone has read the state
            pclk = clock64();
            const bool more = grid_top_rebuild(M, T, Tmin, sh_best, sh_hist);
            g_theta = 0; g_sticky = false; g_cached = 0;
            ML_PHASE(MS_CLK_TOPREB, pclk);
            if (gtid == 0) sh_phase[MS_N_TOPREB - 40]++;
            if (!more) { if (gtid == 0) M.state[MS_DONE] = 1; break; }
            continue;
        }

        // ---- leader mode while the work per merge is small
        if (skip > 0) skip--;
        else if (T2 > 0 && alog_n + ML_LEADER_ITEMS_MAX <= M.alog_cap && !((M.batch_max >> 16) & 1)) {      // bit 16: no leader mode (tuning)
            const i64 m0 = m;
            grid_barrier(M);                                    // everyone has read the state the leader is about to change
            if (blockIdx.x == 0) {
                pclk = clock64();
                leader_loop(M, *(LeaderCtx*)ml_dyn_smem, GB, sh_best, T, Tmin, T2, T2pa);
                ML_PHASE(MS_CLK_LEADER, pclk);
                __syncthreads();
                if (threadIdx.x == 0) { __threadfence(); atomicAdd((u64*)&M.state[MS_LEADER_GEN], 1ULL); }
            } else if (helper_idx >= 0) {
                helper_loop(M, gen, helper_idx, (u64*)sh_hist, &R);        // sh_hist (4 KB) is free while the leader runs
            } else {
                if (threadIdx.x == 0) while (*(volatile i64*)&M.state[MS_LEADER_GEN] == gen) __nanosleep(2000);
                __syncthreads();
            }
            grid_barrier(M);
            m = M.state[MS_NMERGES];
            n_tok = (int32_t)M.state[MS_NTOK];
            alog_n = M.state[MS_ALOG_N];
            const i64 reason = M.state[MS_LEADER_REASON];
            if (m == m0 && reason == LR_OTHER) { backoff = backoff < 64 ? backoff * 2 : 64; skip = backoff; } else backoff = 1;
            if (m >= M.num_merges || M.state[MS_ERROR]) break;
            grid_barrier(M);                                    // everyone has re-read the state
            if (reason == LR_TOP) continue;                 // top list exhausted: rebuild it first
            if (reason == LR_REBUILD) { pclk = clock64(); rebuild_index(M, sh_scan, m); ML_PHASE(MS_CLK_IDXREB, pclk); continue; }
        }

        // ---- one merge in grid mode.  With a valid top list every CTA finds the best pair on its own
        //      (same data, same answer); the barrier only keeps the rewrite (which changes the counts)
        //      from starting before every CTA has read them.
        Best best;
        bool have_best = false;
        pclk = clock64();
        ML_CLOCK(g0);
        // ---- a BATCH of merges in grid mode (rules: "batched leader merges"): every CTA selects the same members from the
        //      same list and counts, builds the same candidate ranges, and the groups of all CTAs share the items.
        //      Barrier 1: every CTA has read the counts.  Barrier 2: every word is rewritten; then every CTA closes the members.
        if (T2 > 0 && T2pa == 0 && M.state[MS_TOP_OVF] == 0 && ((M.batch_max >> 8) & 255 ? (M.batch_max >> 8) & 255 : M.batch_max & 255) != 1) {
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
            const i64 top_now = M.state[MS_TOP_N];              // not the value read above: a leader session may have added entries since
            const int tn = (int)(top_now < ML_TOP_N ? top_now : ML_TOP_N);
            int gbmax = (int)((M.batch_max >> 8) & 255 ? (M.batch_max >> 8) & 255 : M.batch_max & 255);      // bits 8..15: grid mode's own limit (tuning)
            if (gbmax <= 0 || gbmax > ML_BATCH_MAX) gbmax = ML_BATCH_MAX;
            Best mine{0, -1, 0, 0, 0};
            if ((int)threadIdx.x < tn) {
                // the list only grows between rebuilds: every CTA keeps a copy, so the count is ONE round trip away
                if ((int)threadIdx.x >= g_cached) {
                    const u64 k0 = __ldcg(&M.top_key[threadIdx.x]);
                    g_tslot[threadIdx.x] = __ldcg(&M.top_slot[threadIdx.x]); g_tkey[threadIdx.x] = k0;
                    g_tpa[threadIdx.x] = __ldcg(&M.tok_pre[(int32_t)((k0 >> 32) & 0x7fffffff)]);      // tie-breaks without a round trip
                    g_tpb[threadIdx.x] = __ldcg(&M.tok_pre[(int32_t)(k0 & 0xffffffffu)]);
                }
                const int32_t sl = g_tslot[threadIdx.x];
                const u64 k = g_tkey[threadIdx.x];
                const i64 cnt = __ldcg(&M.pcnt[sl]);
                if (cnt > 0) mine = Best{cnt, sl, (int32_t)((k >> 32) & 0x7fffffff), (int32_t)(k & 0xffffffffu), (int32_t)threadIdx.x};
            }
            g_cached = tn;
            int nb = select_batch(M, GB.sel, mine, tn, gbmax, T, Tmin, T2, g_tpa, g_tpb, sh_cnt, g_theta, g_sticky, ML_HEAD);
            long long gclk = pclk;
            ML_PHASE(MS_CLK_GB_SELECT, gclk);
            if (m + nb > M.num_merges) nb = (int)(M.num_merges - m);
            int32_t ma = 0, mb = 0, mslot = -1; i64 mcnt = 0;
            if (lane < ML_BATCH_MAX) { ma = GB.sel.mem[lane].a; mb = GB.sel.mem[lane].b; mslot = GB.sel.mem[lane].slot; mcnt = GB.sel.mem[lane].cnt; }
            if (nb >= 1) {
                best = Best{__shfl_sync(0xffffffffu, mcnt, 0), __shfl_sync(0xffffffffu, mslot, 0), __shfl_sync(0xffffffffu, ma, 0), __shfl_sync(0xffffffffu, mb, 0), 0};
                have_best = true;
            }
            if (nb >= 1) {                                      // (a "batch" of one included: the rewrite below has the shorter chain)
                grid_barrier(M);                                // 1: every CTA has read the counts
                ML_PHASE(MS_CLK_GB_BAR1, gclk);
                if (warp < nb && lane == warp) build_ranges(M, mslot, ma, mb, &GB.R[warp]);
                if (nwarps - 1 - warp < nb && lane == nwarps - 1 - warp) GB.c[lane] = lookup_merged(M, ma, mb, n_tok, &GB.MI[lane]);
                if (warp == nwarps / 2 && lane < nb) GB.mkey[lane] = PAIR_KEY(ma, mb);
                __syncthreads();
                int mtot = 0, mnew = 0, mlen = 0; int32_t mc = -1; u64 mH = 0; bool mbad = true;
                if (lane < nb) {
                    mtot = (int)(GB.R[lane].total < 0x08000000 ? GB.R[lane].total : 0x08000000);
                    mbad = GB.R[lane].n < 0;
                    mc = GB.c[lane]; mnew = mc == n_tok ? 1 : 0; mH = GB.MI[lane].H; mlen = (int)(GB.MI[lane].la + GB.MI[lane].lb);
                    if (mnew) mc = n_tok + lane;
                }
                int cum = mtot, lcum = mlen;
#pragma unroll
                for (int o = 1; o < ML_BATCH_MAX; o <<= 1) {
                    const int t1 = __shfl_up_sync(0xffffffffu, cum, o), t3 = __shfl_up_sync(0xffffffffu, lcum, o);
                    if (lane >= o) { cum += t1; lcum += t3; }
                }
                if (alog_n + cum > M.alog_cap) mbad = true;
                for (int i = 0; i < nb - 1; i++) {
                    const u64 Hi = __shfl_sync(0xffffffffu, mH, i);
                    const int newi = __shfl_sync(0xffffffffu, mnew, i);
                    if (i < lane && (Hi == mH || !newi)) mbad = true;
                }
                const int kk = __ffs(__ballot_sync(0xffffffffu, mbad)) - 1;
                __syncthreads();                                // GB is rewritten by the next selection
                ML_PHASE(MS_CLK_GB_RANGES, gclk);
                const int warps_all = (int)gridDim.x * nwarps;
                if (kk >= 2 || (kk == 1 && __shfl_sync(0xffffffffu, cum, 0) <= warps_all * 8)) {      // heavy single merges: thread per word, below
                    const int items_all = __shfl_sync(0xffffffffu, cum, kk - 1);
                    const int G = items_all > (warps_all - 1) * 8 ? 2 : (items_all > (warps_all - 1) * 4 ? 4 : 8), padm = 32 / G - 1;
                    int pcum = lane < kk ? (mtot + padm) & ~padm : 0;
#pragma unroll
                    for (int o = 1; o < ML_BATCH_MAX; o <<= 1) { const int t2 = __shfl_up_sync(0xffffffffu, pcum, o); if (lane >= o) pcum += t2; }
                    const int mpst = pcum - (lane < kk ? (mtot + padm) & ~padm : 0), mseg = (int)alog_n + cum - mtot;
                    const i64 pool_end = M.tok_off[n_tok];
                    const i64 moc = pool_end + lcum - mlen;
                    const int last_new = __shfl_sync(0xffffffffu, mnew, kk - 1) ? kk - 1 : kk - 2;
                    if (blockIdx.x == 0) {
                        if (warp == nwarps - 1 && lane < kk) { commit_member(M, nullptr, m + lane, ma, mb, mc, mnew != 0, (i64)mseg, GB.MI[lane], moc, lane == last_new); M.pcnt[mslot] = 0; }
                    }
                    const int32_t stamp = (int32_t)(m + 1);
                    // the committing warp (last of CTA 0) takes no items: its round trips would sit on every batch's critical path
                    // consecutive items go to warps of DIFFERENT CTAs (a batch of 1 000 items keeps 148 SMs busy, not 8)
                    const int vw = warp * (int)gridDim.x + (int)blockIdx.x, skipw = (nwarps - 1) * (int)gridDim.x;
                    const int gw = vw - (vw > skipw ? 1 : 0), rw = warps_all - 1;
                    if (blockIdx.x == 0 && warp == nwarps - 1) { }
                    else if (G == 8) rewrite_batch<8>(M, nullptr, GB.R, GB.mkey, kk, lane, gw * 4, rw * 4, stamp, ma, mb, mc, mslot, mnew, mtot, mpst, mseg, T, T2);
                    else if (G == 4) rewrite_batch<4>(M, nullptr, GB.R, GB.mkey, kk, lane, gw * 8, rw * 8, stamp, ma, mb, mc, mslot, mnew, mtot, mpst, mseg, T, T2);
                    else rewrite_batch<2>(M, nullptr, GB.R, GB.mkey, kk, lane, gw * 16, rw * 16, stamp, ma, mb, mc, mslot, mnew, mtot, mpst, mseg, T, T2);
                    ML_PHASE(MS_CLK_GB_REWRITE, gclk);
                    grid_barrier(M);                            // 2: every word is rewritten, every token created
                    ML_PHASE(MS_CLK_GB_BAR2, gclk);
                    // every CTA writes the same values: no further barrier before the next iteration
                    if (warp == 0 && lane < kk) {
                        const int e = mseg + mtot;
                        M.seg_end[m + lane] = e; M.tok_first[mc] = (int32_t)(m + lane);
                        M.tok_head[mc] = make_int4((int32_t)(m + lane), mseg, e, M.merge_next[m + lane]);
                    }
                    if (threadIdx.x == 0) { M.state[MS_NMERGES] = m + kk; M.state[MS_ALOG_N] = alog_n + items_all; }
                    if (gtid == 0) {
                        M.state[MS_GRID_MERGES] += kk;
                        if (kk >= 2) { sh_phase[MS_GRID_ITERS - 40]++; sh_phase[MS_GRID_BATCHED - 40] += kk; }
                        else { const int cls = (i64)items_all * 64 <= gstride ? 0 : 1; sh_phase[MS_GRID_CLS + cls - 40]++; sh_phase[MS_GRID_CLS + 3 + cls - 40] += clock64() - pclk; }
                    }
                    ML_PHASE(MS_CLK_GRID, pclk);
                    __syncthreads();
                    continue;
                }
                // fewer than two members fit (candidate index / log space, product tokens): the best pair goes alone
            }
        }
        if (T2 > 0) {
            if (!have_best) best = top_best(M, M.state[MS_TOP_N], sh_best, sh_cnt);
            grid_barrier(M);
            if (best.slot < 0 || best.cnt < T2 || best.cnt < T || (best.cnt == T2 && T2pa != 0 && M.tok_pre[best.a] < T2pa)) {
                if (gtid == 0) M.state[MS_T2] = 0;
                grid_barrier(M);
                continue;
            }
        } else {
            // top list disabled (massive ties): full scan of the active set for this merge
            best = grid_argmax(M, sh_best);
            if (best.slot < 0 || best.cnt < T) {            // cannot happen right after a rebuild; be safe
                grid_barrier(M);
                if (gtid == 0) M.state[MS_T2] = 0;
                grid_barrier(M);
                continue;
            }
        }
        const int32_t a = best.a, b = best.b;
        ML_CLOCK(g1);
        if (threadIdx.x == 0) { build_ranges(M, best.slot, a, b, &R); sh_c = lookup_merged(M, a, b, n_tok); }
        __syncthreads();
        ML_CLOCK(g2);
        if (R.n < 0 || alog_n + R.total > M.alog_cap) {
            // fold the affected log into the CSR index first, then look the candidates up again
            grid_barrier(M);
            rebuild_index(M, sh_scan, m);
            alog_n = 0;
            if (threadIdx.x == 0) build_ranges(M, best.slot, a, b, &R);
            __syncthreads();
        }
        const int32_t c = sh_c;
        const bool is_new = c == n_tok;
        if (blockIdx.x == 0) { commit_merge(M, m, a, b, c, is_new, alog_n); if (threadIdx.x == 0) M.pcnt[best.slot] = 0; }
        const int32_t stamp = (int32_t)(m + 1);
        const i64 T2u = T2 > 0 ? T2 : 0;
        if (a != b && R.total * 8 <= gstride) {
            // up to one candidate per 8-lane group in a single pass: the stamp exchange and the word header loads of
            // a candidate are issued together (one DRAM round trip instead of two), four words per warp in flight
            const i64 gg = gtid >> 3;
            const int lane = threadIdx.x & 31, gl = lane & 7;
            int32_t w = -1;
            i64 off = 0, f = 0; int n = 0;
            int32_t old = stamp;
            if (gg < R.total) {
                const i64 e = range_item(R, gg);
                w = POST_WORD(e); off = POST_OFF(e);
                if (gl == 0) { old = atomicExch(&M.wstamp[w], stamp); prefetch_l2(&M.wsym[off]); prefetch_l2(&M.wslot[off]); }
                n = M.wlen[w]; f = M.wcnt[w];
            }
            old = __shfl_sync(0xffffffffu, old, lane & 24);       // all 32 lanes: groups beyond the list carry old == stamp
            if (old == stamp) w = -1;
            rewrite_words_g<8>(M, w, off, n, f, a, b, c, T, T2u, nullptr, is_new);
        } else if (a != b && R.total * 32 <= gstride * 4) {
            const i64 gw = gtid >> 5, nw = gstride >> 5;
            for (i64 it = gw; it < R.total; it += nw) {
                const int32_t w = POST_WORD(range_item(R, it));
                int32_t old = 0;
                if ((threadIdx.x & 31) == 0) old = atomicExch(&M.wstamp[w], stamp);
                old = __shfl_sync(0xffffffffu, old, 0);
                if (old == stamp) continue;
                rewrite_word_warp(M, w, a, b, c, T, T2u, nullptr, is_new);
            }
        } else {
            for (i64 it = gtid; it < R.total; it += gstride) {
                const int32_t w = POST_WORD(range_item(R, it));
                if (atomicExch(&M.wstamp[w], stamp) == stamp) continue;
                rewrite_word_thread(M, w, a, b, c, T, T2u, nullptr, is_new, best.slot);
            }
        }
        ML_CLOCK(g3);
        grid_barrier(M);
        ML_CLOCK(g4);
#if ML_TIMING
        if (gtid == 0) { M.state[12] += g1 - g0; M.state[13] += g2 - g1; M.state[22] += g3 - g2; M.state[27] += g4 - g3; M.state[28] += R.total; }
#endif
        // every CTA writes the same values: no further barrier needed before the next iteration
        // Massive ties (T2 < 0: every merge scans the active set): rebuilding the top list after EVERY merge only to see it
        // overflow again cost ten grid barriers per merge; try again when the maximum count has moved or after 64 merges.
        if (threadIdx.x == 0) {
            close_merge(M, m, c);
            if (T2 < 0) { if (best.cnt != tie_cnt || tie_left <= 1) M.state[MS_T2] = 0; else M.state[MS_TIE_LEFT] = tie_left - 1; }
        }
        if (gtid == 0) {
            M.state[MS_GRID_MERGES]++;
            const int cls = R.total * 64 <= gstride ? 0 : (R.total * 8 <= gstride ? 1 : 2);
            sh_phase[MS_GRID_CLS + cls - 40]++; sh_phase[MS_GRID_CLS + 3 + cls - 40] += clock64() - pclk;
        }
        ML_PHASE(MS_CLK_GRID, pclk);
        __syncthreads();
    }
    if (gtid == 0) {
#if ML_RW_TRACE
        for (int i = 0; i < 6; i++) { sh_phase[8 + i] = (long long)g_rw_clk[(ML_RW_TRACE == 2 ? 8 : 0) + i]; g_rw_clk[i] = 0; g_rw_clk[8 + i] = 0; }
#endif
        for (int i = 0; i < 24; i++) if (sh_phase[i]) M.state[40 + i] += sh_phase[i];
        M.state[MS_CLK_TOTAL] = clock64() - pclk0;
    }
}
