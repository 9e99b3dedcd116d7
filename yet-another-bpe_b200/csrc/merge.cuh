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
// (pairs with count >= T) -> grid sync -> rewrite of the words listed in the pair's postings
// (CSR index + delta log) with incremental pair-count deltas -> grid sync.
#pragma once

#include <cooperative_groups.h>

#include "pretok.cuh"

namespace cg = cooperative_groups;

#define ML_THREADS 512
#define PAIR_KEY(a, b) (0x8000000000000000ULL | ((u64)(uint32_t)(a) << 32) | (u64)(uint32_t)(b))
#define TOK_HASH_B 0x100000001b3ULL

// merge-loop state slots (device int64[32])
#define MS_NMERGES 0
#define MS_NTOK 1
#define MS_ERROR 2
#define MS_DLOG_N 3
#define MS_ACT_N 4
#define MS_T 5
#define MS_NPAIRS 6
#define MS_DLOG_OVF 7
#define MS_POOL_USED 8
#define MS_REBUILDS 9
#define MS_TREBUILDS 10
#define MS_MAXCNT 11
#define MS_SCRATCH 12

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

__global__ void __launch_bounds__(256) k_compact_short(const ulonglong2* keys, const i64* counts, i64 cap, WordTable W) {
    i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += stride) {
        ulonglong2 kv = keys[i];
        int32_t wid = -1;
        if (kv.y != 0) {
            int len = (int)(kv.x >> 56);
            wid = (int32_t)atomicAdd((u64*)&W.counters[0], 1ULL);
            i64 off = (i64)atomicAdd((u64*)&W.counters[1], (u64)len);
            W.woff[wid] = off; W.wlen[wid] = len; W.wcnt[wid] = counts[i];
            for (int k = 0; k < len; k++) {
                int b = k < 7 ? (int)((kv.x >> (8 * k)) & 0xff) : (int)((kv.y >> (8 * (k - 7))) & 0xff);
                W.wsym[off + k] = b; W.sym_word[off + k] = wid;
            }
        }
        if (W.sword) W.sword[i] = wid;
    }
}

__global__ void __launch_bounds__(256) k_compact_long(const LongEntry* ent, i64 cap, const uint8_t* text, WordTable W) {
    __shared__ i64 sh_off; __shared__ int32_t sh_wid;
    for (i64 i = blockIdx.x; i < cap; i += gridDim.x) {
        LongEntry e = ent[i];
        bool occ = e.h >= 2;
        if (threadIdx.x == 0) {
            sh_wid = -1;
            if (occ) {
                sh_wid = (int32_t)atomicAdd((u64*)&W.counters[0], 1ULL);
                sh_off = (i64)atomicAdd((u64*)&W.counters[1], (u64)e.len);
                W.woff[sh_wid] = sh_off; W.wlen[sh_wid] = (int32_t)e.len; W.wcnt[sh_wid] = e.count;
            }
            if (W.lword) W.lword[i] = sh_wid;
        }
        __syncthreads();
        if (occ) {
            i64 off = sh_off; int32_t wid = sh_wid;
            for (i64 k = threadIdx.x; k < e.len; k += blockDim.x) { W.wsym[off + k] = text[e.pos + k]; W.sym_word[off + k] = wid; }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------
// merge loop
// ---------------------------------------------------------------------------------
struct Best { i64 cnt; int32_t slot; int32_t a; int32_t b; int32_t pad; };

struct MergeParams {
    // words
    int32_t* wsym; const int32_t* sym_word; i64 n_syms;
    const i64* woff; int32_t* wlen; const i64* wcnt; i64 n_words; int32_t* wstamp;
    // tokens (ids < n_base are prepared by the host: 256 bytes + specials)
    uint8_t* tok_bytes; i64 tok_bytes_cap; i64* tok_off; u64* tok_hash; u64* tok_pow;
    u64* tset; i64 tset_cap; i64 max_tokens;
    // pairs
    u64* pkey; i64* pcnt; i64 pcap;
    uint32_t* ioff; uint32_t* icnt; int32_t* ipost; uint32_t* inact; int32_t* act;
    int32_t* dlog_slot; int32_t* dlog_word; i64 dlog_cap;
    Best* partial; i64* bsum;
    // outputs
    int32_t* merges; int32_t* merge_new; i64* state;
    i64 num_merges; i64 min_freq;
};

__device__ __forceinline__ i64 pair_find(const MergeParams& M, u64 key) {
    u64 mask = (u64)M.pcap - 1;
    u64 slot = mix64(key) & mask;
    for (i64 probes = 0; probes < M.pcap; probes++) {
        u64 k = M.pkey[slot];
        if (k == key) return (i64)slot;
        if (k == 0) return -1;
        slot = (slot + 1) & mask;
    }
    return -1;
}
__device__ __forceinline__ i64 pair_upsert(const MergeParams& M, u64 key) {
    u64 mask = (u64)M.pcap - 1;
    u64 slot = mix64(key) & mask;
    for (i64 probes = 0; probes < M.pcap; probes++) {
        u64 k = *(volatile u64*)&M.pkey[slot];
        if (k == 0) {
            k = atomicCAS(&M.pkey[slot], 0ULL, key);
            if (k == 0) { atomicAdd((u64*)&M.state[MS_NPAIRS], 1ULL); return (i64)slot; }
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

// grid-wide argmax over the active set; every block returns the same result
__device__ Best grid_argmax(const MergeParams& M, cg::grid_group& grid, Best* sh) {
    i64 n = M.state[MS_ACT_N];
    Best v{0, -1, 0, 0, 0};
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        int32_t slot = M.act[i];
        i64 c = M.pcnt[slot];
        if (c <= 0 || c < v.cnt) continue;
        u64 k = M.pkey[slot];
        Best t{c, slot, (int32_t)((k >> 32) & 0x7fffffff), (int32_t)(k & 0xffffffffu), 0};
        if (best_gt(M, t, v)) v = t;
    }
    v = block_best(M, v, sh);
    if (threadIdx.x == 0) M.partial[blockIdx.x] = v;
    grid.sync();
    Best w{0, -1, 0, 0, 0};
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) { Best t = M.partial[i]; if (best_gt(M, t, w)) w = t; }
    return block_best(M, w, sh);
}

__device__ void rebuild_active(const MergeParams& M, cg::grid_group& grid, i64 T) {
    i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gstride = (i64)gridDim.x * blockDim.x;
    for (i64 i = gtid; i < (M.pcap + 31) / 32; i += gstride) M.inact[i] = 0;
    if (gtid == 0) { M.state[MS_ACT_N] = 0; M.state[MS_TREBUILDS]++; }
    grid.sync();
    for (i64 s = gtid; s < M.pcap; s += gstride) {
        if (M.pkey[s] != 0 && M.pcnt[s] >= T) {
            atomicOr(&M.inact[s >> 5], 1u << (s & 31));
            i64 idx = (i64)atomicAdd((u64*)&M.state[MS_ACT_N], 1ULL);
            M.act[idx] = (int32_t)s;
        }
    }
    grid.sync();
}

// CSR postings: for every pair slot the words that contain it (duplicates allowed)
__device__ void rebuild_index(const MergeParams& M, cg::grid_group& grid, i64* sh_scan) {
    i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gstride = (i64)gridDim.x * blockDim.x;
    for (i64 i = gtid; i < M.pcap; i += gstride) M.icnt[i] = 0;
    grid.sync();
    for (i64 i = gtid; i + 1 < M.n_syms; i += gstride) {
        int32_t w = M.sym_word[i];
        i64 j = i - M.woff[w];
        if (j + 1 < M.wlen[w]) {
            i64 s = pair_find(M, PAIR_KEY(M.wsym[i], M.wsym[i + 1]));
            if (s >= 0) atomicAdd(&M.icnt[s], 1u);
            else atomicOr((u64*)&M.state[MS_ERROR], (u64)ME_INTERNAL);
        }
    }
    grid.sync();
    // exclusive scan of icnt -> ioff, one contiguous chunk per block
    i64 chunk = (M.pcap + gridDim.x - 1) / gridDim.x;
    i64 lo = chunk * blockIdx.x, hi = lo + chunk < M.pcap ? lo + chunk : M.pcap;
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
    grid.sync();
    {
        i64 base = 0;
        for (int b = 0; b < (int)blockIdx.x; b++) base += M.bsum[b];
        __syncthreads();
        for (i64 t0 = lo; t0 < hi; t0 += blockDim.x) {
            i64 i = t0 + threadIdx.x;
            i64 v = i < hi ? M.icnt[i] : 0;
            // block inclusive scan
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
        if (hi == M.pcap && lo < hi && blockIdx.x != gridDim.x - 1 && threadIdx.x == 0) M.ioff[M.pcap] = (uint32_t)base;
    }
    grid.sync();
    for (i64 i = gtid; i + 1 < M.n_syms; i += gstride) {
        int32_t w = M.sym_word[i];
        i64 j = i - M.woff[w];
        if (j + 1 < M.wlen[w]) {
            i64 s = pair_find(M, PAIR_KEY(M.wsym[i], M.wsym[i + 1]));
            if (s >= 0) { uint32_t r = atomicSub(&M.icnt[s], 1u) - 1; M.ipost[M.ioff[s] + r] = w; }
        }
    }
    if (gtid == 0) { M.state[MS_DLOG_N] = 0; M.state[MS_DLOG_OVF] = 0; M.state[MS_REBUILDS]++; }
    grid.sync();
}

__device__ __forceinline__ void pair_sub(const MergeParams& M, int32_t x, int32_t y, i64 f) {
    i64 s = pair_find(M, PAIR_KEY(x, y));
    if (s >= 0) atomicAdd((u64*)&M.pcnt[s], (u64)(-f));
    else atomicOr((u64*)&M.state[MS_ERROR], (u64)ME_INTERNAL);
}
__device__ __forceinline__ void pair_add(const MergeParams& M, int32_t x, int32_t y, i64 f, int32_t w, i64 T) {
    i64 s = pair_upsert(M, PAIR_KEY(x, y));
    if (s < 0) return;
    i64 now = (i64)atomicAdd((u64*)&M.pcnt[s], (u64)f) + f;
    if (now >= T) {
        uint32_t bit = 1u << (s & 31);
        if (!(atomicOr(&M.inact[s >> 5], bit) & bit)) {
            i64 idx = (i64)atomicAdd((u64*)&M.state[MS_ACT_N], 1ULL);
            M.act[idx] = (int32_t)s;
        }
    }
    i64 d = (i64)atomicAdd((u64*)&M.state[MS_DLOG_N], 1ULL);
    if (d < M.dlog_cap) { M.dlog_slot[d] = (int32_t)s; M.dlog_word[d] = w; }
    else M.state[MS_DLOG_OVF] = 1;
}

// rewrite one word in place (left->right, non-overlapping) and apply the pair-count deltas
__device__ void rewrite_word(const MergeParams& M, int32_t w, int32_t a, int32_t b, int32_t c, i64 T) {
    int32_t* s = M.wsym + M.woff[w];
    int n = M.wlen[w];
    i64 f = M.wcnt[w];
    int o = 0, j = 0;
    int32_t prev_old = -1, prev_new = -1;
    bool prev_changed = false;
    while (j < n) {
        int32_t x = s[j];
        if (j + 1 < n && x == a && s[j + 1] == b) {
            if (o > 0) { pair_sub(M, prev_old, a, f); pair_add(M, prev_new, c, f, w, T); }
            pair_sub(M, a, b, f);
            s[o++] = c; prev_old = b; prev_new = c; prev_changed = true; j += 2;
        } else {
            if (o > 0 && prev_changed) { pair_sub(M, prev_old, x, f); pair_add(M, prev_new, x, f, w, T); }
            s[o++] = x; prev_old = x; prev_new = x; prev_changed = false; j += 1;
        }
    }
    M.wlen[w] = o;
}

__global__ void __launch_bounds__(ML_THREADS) k_merge_loop(MergeParams M) {
    cg::grid_group grid = cg::this_grid();
    __shared__ Best sh_best[ML_THREADS / 32];
    __shared__ i64 sh_scan[1 + ML_THREADS / 32];
    __shared__ int32_t sh_c;
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gstride = (i64)gridDim.x * blockDim.x;

    // ---- K4: pair histogram (trainer.py:228-235)
    for (i64 i = gtid; i + 1 < M.n_syms; i += gstride) {
        int32_t w = M.sym_word[i];
        i64 j = i - M.woff[w];
        if (j + 1 < M.wlen[w]) {
            i64 s = pair_upsert(M, PAIR_KEY(M.wsym[i], M.wsym[i + 1]));
            if (s >= 0) atomicAdd((u64*)&M.pcnt[s], (u64)M.wcnt[w]);
        }
    }
    grid.sync();
    // global max count -> first threshold
    {
        i64 mx = 0;
        for (i64 s = gtid; s < M.pcap; s += gstride) { i64 c = M.pcnt[s]; if (c > mx) mx = c; }
        for (int o = 16; o > 0; o >>= 1) { i64 t = __shfl_xor_sync(0xffffffffu, mx, o); if (t > mx) mx = t; }
        if ((threadIdx.x & 31) == 0 && mx > 0) atomicMax((i64*)&M.state[MS_MAXCNT], mx);
    }
    rebuild_index(M, grid, sh_scan);      // starts and ends with grid-wide syncs
    i64 Tmin = M.min_freq > 1 ? M.min_freq : 1;
    i64 T = M.state[MS_MAXCNT] / 4; if (T < Tmin) T = Tmin;
    rebuild_active(M, grid, T);

    int32_t n_tok = (int32_t)M.state[MS_NTOK];   // every block tracks the token count identically
    i64 m = 0;
    for (; m < M.num_merges; m++) {
        if (M.state[MS_ERROR]) break;             // uniform: read right after a grid sync
        // ---- phase 1: best pair
        Best best = grid_argmax(M, grid, sh_best);
        bool stop = false;
        while (best.slot < 0 || best.cnt < T) {
            if (T <= Tmin) { stop = true; break; }   // nothing left with count >= max(1, min_frequency)
            T = T / 4; if (T < Tmin) T = Tmin;
            grid.sync();                            // everyone has read partial[] / act before it is rebuilt
            rebuild_active(M, grid, T);
            best = grid_argmax(M, grid, sh_best);
        }
        if (stop) break;
        const int32_t a = best.a, b = best.b;

        // ---- phase 2: merged token id (existing id when the bytes are already a token)
        if (threadIdx.x == 0) {
            u64 H = M.tok_hash[a] * M.tok_pow[b] + M.tok_hash[b];
            i64 la = M.tok_off[a + 1] - M.tok_off[a], lb = M.tok_off[b + 1] - M.tok_off[b];
            int32_t c = n_tok;
            u64 mask = (u64)M.tset_cap - 1, slot = mix64(H) & mask;
            for (;;) {
                u64 e = *(volatile u64*)&M.tset[slot];
                if (e == 0) break;
                int32_t id = (int32_t)(e & 0xffffffffu) - 1;
                if ((e >> 32) == (H >> 32) && id < n_tok && M.tok_hash[id] == H && M.tok_off[id + 1] - M.tok_off[id] == la + lb) {
                    const uint8_t* pc = M.tok_bytes + M.tok_off[id];
                    const uint8_t* pa = M.tok_bytes + M.tok_off[a];
                    const uint8_t* pb = M.tok_bytes + M.tok_off[b];
                    bool eq = true;
                    for (i64 k = 0; k < la && eq; k++) eq = pc[k] == pa[k];
                    for (i64 k = 0; k < lb && eq; k++) eq = pc[la + k] == pb[k];
                    if (eq) { c = id; break; }
                }
                slot = (slot + 1) & mask;
            }
            sh_c = c;
        }
        __syncthreads();
        const int32_t c = sh_c;
        const bool is_new = c == n_tok;
        if (blockIdx.x == 0) {
            if (threadIdx.x == 0) { M.merges[2 * m] = a; M.merges[2 * m + 1] = b; M.merge_new[m] = c; }
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
                        __threadfence();
                        u64 mask = (u64)M.tset_cap - 1, slot = mix64(H) & mask;
                        while (M.tset[slot] != 0) slot = (slot + 1) & mask;
                        atomicExch(&M.tset[slot], (H & 0xffffffff00000000ULL) | (u64)(uint32_t)(c + 1));
                        M.state[MS_NTOK] = c + 1; M.state[MS_POOL_USED] = oc + la + lb;
                    }
                }
            }
            if (threadIdx.x == 0) M.state[MS_NMERGES] = m + 1;
        }
        if (is_new) n_tok++;

        // ---- phase 3: rewrite the words that contain (a, b)
        {
            const i64 p0 = M.ioff[best.slot], npost = (i64)M.ioff[best.slot + 1] - p0;
            i64 nlog = M.state[MS_DLOG_N]; if (nlog > M.dlog_cap) nlog = M.dlog_cap;
            grid.sync();     // snapshot of dlog_n taken by every block before anyone appends
            const int32_t stamp = (int32_t)(m + 1);
            for (i64 it = gtid; it < npost + nlog; it += gstride) {
                int32_t w;
                if (it < npost) w = M.ipost[p0 + it];
                else { i64 d = it - npost; w = M.dlog_slot[d] == best.slot ? M.dlog_word[d] : -1; }
                if (w < 0) continue;
                if (atomicExch(&M.wstamp[w], stamp) == stamp) continue;
                rewrite_word(M, w, a, b, c, T);
            }
        }
        grid.sync();
        // ---- maintenance: fold the delta log into the CSR index when it fills up
        if (M.state[MS_DLOG_OVF] || M.state[MS_DLOG_N] > M.dlog_cap / 2) {
            grid.sync();
            rebuild_index(M, grid, sh_scan);
        }
    }
    (void)m;
}
