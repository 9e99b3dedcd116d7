// encode_small.cuh -- one-launch encode for short inputs (<= ES_MAX_BYTES).
//
// Replaces /root/reference/src/yet_another_bpe/tokenizer.py:152-308 for the strings the reference's own tests pass to
// `encode` (a few bytes to a few KB).  The batched pipeline (encode.cuh) costs 13 launches and three host round trips
// whatever the input length; here ONE CTA does everything and the host waits for one event:
//   text (device or MAPPED pinned host memory) -> shared memory
//   special-token candidates (thread per byte), leftmost / longest-first resolution (tokenizer.py:97-102,171)
//   pre-token start bit per byte: the generic rule of common.cuh (is_token_start), the same code k_token_starts uses
//   pre-tokens dealt out evenly over the threads (one each up to 1 024 of them): BPE by rank in thread-local memory
//     (es_encode_token: the heap loop of tokenizer.py:247-294 with ~3 n table probes), ids compacted into the scratch
//     range the thread's tokens cover
//   block scan of the per-thread id counts -> ids in text order -> out[1..], out[0] = count (or -1: fall back)
// Pre-tokens longer than ES_MAX_TOKEN bytes (the thread-per-token BPE is quadratic) make the kernel report -1; the caller
// then takes the batched path.  No hash tables, no de-duplication: a short text has few repeats.
#pragma once

#include "encode.cuh"

#define ES_MAX_BYTES 32768
#define ES_THREADS 1024
#define ES_MAX_TOKEN 64
#define ES_WORDS (ES_MAX_BYTES / 32)

__device__ __forceinline__ int es_next_start(const uint32_t* sbits, int from, int n) {      // first start > from, or n
    int i = from + 1;
    while (i < n) {
        const uint32_t w = sbits[i >> 5] >> (i & 31);
        if (w) { const int r = i + __ffs(w) - 1; return r < n ? r : n; }
        i = ((i >> 5) + 1) << 5;
    }
    return n;
}

// tokenizer.py:247-294 on one pre-token in thread-local memory: merge the adjacent pair with the smallest (rank, position),
// one occurrence at a time.  The rank of every adjacency is looked up once (independent loads) and only the two adjacencies
// next to a merge are looked up again: about 3 n table probes per pre-token instead of n^2 / 2.
__device__ int es_encode_token(const EncodeModel& E, int32_t* sym, int n) {
    uint32_t rk[ES_MAX_TOKEN]; int32_t rs[ES_MAX_TOKEN];
    for (int j = 0; j < n; j++) sym[j] = E.byte_sym[sym[j]];
    for (int j = 0; j + 1 < n; j++) { uint32_t r; int32_t res; rk[j] = merge_rank(E, sym[j], sym[j + 1], &r, &res) ? r : 0xffffffffu; rs[j] = res; }
    while (n > 1) {
        uint32_t br = 0xffffffffu; int bp = -1;
        for (int j = 0; j + 1 < n; j++) if (rk[j] < br) { br = rk[j]; bp = j; }
        if (bp < 0) break;
        sym[bp] = rs[bp];
        for (int j = bp + 1; j + 1 < n; j++) { sym[j] = sym[j + 1]; rk[j] = rk[j + 1]; rs[j] = rs[j + 1]; }
        n--;
        uint32_t r; int32_t res;
        if (bp > 0) { rk[bp - 1] = merge_rank(E, sym[bp - 1], sym[bp], &r, &res) ? r : 0xffffffffu; rs[bp - 1] = res; }
        if (bp + 1 < n) { rk[bp] = merge_rank(E, sym[bp], sym[bp + 1], &r, &res) ? r : 0xffffffffu; rs[bp] = res; }
    }
    return n;
}

__global__ void __launch_bounds__(ES_THREADS, 1) k_encode_small(EncodeModel E, const uint8_t* __restrict__ text, int n, int n_sp,
                                                                int32_t* __restrict__ scratch, volatile int32_t* out, int out_cap) {
    extern __shared__ __align__(16) unsigned char es_smem[];
    const int nwords = (n + 31) >> 5;
    const int txt_bytes = ((n + 15) & ~15) + 64;
    uint8_t* stxt = es_smem;                                             // n bytes + zero padding
    uint32_t* cand = (uint32_t*)(es_smem + txt_bytes);                   // nwords + 2 each
    uint32_t* rec = cand + nwords + 2;
    uint32_t* sbits = rec + nwords + 2;
    int* wpre = (int*)(sbits + nwords + 2);                              // exclusive prefix of the start counts per 32-byte segment
    __shared__ int sh_scan[ES_THREADS / 32];
    __shared__ int sh_fallback, sh_ntok;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    for (int i = tid; i < txt_bytes / 4; i += ES_THREADS) {
        const int b = 4 * i;
        uint32_t v = 0;
        if (b + 4 <= n && (((uintptr_t)text) & 3) == 0) v = ((const uint32_t*)text)[i];
        else for (int k = 0; k < 4; k++) if (b + k < n) v |= (uint32_t)text[b + k] << (8 * k);
        ((uint32_t*)stxt)[i] = v;
    }
    for (int i = tid; i < 3 * (nwords + 2); i += ES_THREADS) cand[i] = 0;
    if (tid == 0) sh_fallback = 0;
    __syncthreads();

    // ---- specials: candidates, then leftmost-first / priority-order resolution (encode mode: unconditional)
    if (n_sp > 0) {
        for (int base = 0; base < nwords * 32; base += ES_THREADS) {
            const int p = base + tid;
            const bool c = p < n && special_match(stxt, p, n) >= 0;
            const uint32_t m = __ballot_sync(0xffffffffu, c);
            if (lane == 0 && m && (p >> 5) < nwords) cand[p >> 5] = m;
        }
        __syncthreads();
        if (tid == 0) {
            int e = 0;
            for (int w = 0; w < nwords; w++) {
                uint32_t m = cand[w];
                while (m) {
                    const int q = (w << 5) + __ffs(m) - 1; m &= m - 1;
                    if (q < e) continue;
                    const int sp = special_match(stxt, q, n);
                    rec[q >> 5] |= 1u << (q & 31);
                    e = q + (c_sp.offs[sp + 1] - c_sp.offs[sp]);
                }
            }
        }
        __syncthreads();
    }

    // ---- pre-token starts (generic rule over the shared-memory copy)
    {
        GlobalText G{stxt, (i64)n, nullptr, 0, n_sp > 0 ? rec : nullptr, -1, 1};
        for (int base = 0; base < nwords * 32; base += ES_THREADS) {
            const int p = base + tid;
            const bool st = p < n && is_token_start(G, p);
            const uint32_t m = __ballot_sync(0xffffffffu, st);
            if (lane == 0 && (p >> 5) < nwords) sbits[p >> 5] = m;
        }
    }
    __syncthreads();

    // ---- pre-tokens are dealt out evenly: thread t takes tokens [t K, t K + K), K = ceil(T / threads)
    {
        const int c = tid < nwords ? __popc(sbits[tid]) : 0;
        int inc = c;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) sh_scan[wid] = inc;
        __syncthreads();
        int wbase = 0, tot = 0;
        for (int k = 0; k < ES_THREADS / 32; k++) { const int t = sh_scan[k]; if (k < wid) wbase += t; tot += t; }
        if (tid < nwords) wpre[tid] = wbase + inc - c;
        if (tid == 0) { sh_ntok = tot; wpre[nwords] = tot; }
        __syncthreads();
    }
    const int T = sh_ntok, K = (T + ES_THREADS - 1) / ES_THREADS;
    int my_total = 0, my_first = -1;
    if (K > 0 && tid * K < T) {
        const int t0 = tid * K, t1 = t0 + K < T ? t0 + K : T;
        int lo = 0, hi = nwords;                                         // segment of token t0: last w with wpre[w] <= t0
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (wpre[mid] <= t0) lo = mid; else hi = mid; }
        uint32_t m = sbits[lo];
        for (int skip = t0 - wpre[lo]; skip > 0; skip--) m &= m - 1;
        int s = (lo << 5) + __ffs(m) - 1;
        int cursor = s; my_first = s;
        for (int t = t0; t < t1; t++) {
            const int e = es_next_start(sbits, s, n), len = e - s;
            if (n_sp > 0 && ((rec[s >> 5] >> (s & 31)) & 1u)) {          // a special: its id, or dropped (tokenizer.py:177-181)
                const int sp = special_match(stxt, s, n);
                const int32_t id = sp >= 0 ? E.sp_ids[sp] : -1;
                if (id >= 0) scratch[cursor++] = id;
            } else if (len > ES_MAX_TOKEN) sh_fallback = 1;
            else {
                int32_t sym[ES_MAX_TOKEN];
                for (int j = 0; j < len; j++) sym[j] = stxt[s + j];
                const int cnt = es_encode_token(E, sym, len);
                for (int j = 0; j < cnt; j++) scratch[cursor++] = E.sym_out[sym[j]];
            }
            s = e;
        }
        my_total = cursor - my_first;
    }
    // ---- ids in text order
    int inc = my_total;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    __syncthreads();
    if (lane == 31) sh_scan[wid] = inc;
    __syncthreads();
    int wbase = 0, total = 0;
    for (int k = 0; k < ES_THREADS / 32; k++) { const int t = sh_scan[k]; if (k < wid) wbase += t; total += t; }
    const int off = wbase + inc - my_total;
    const bool bad = sh_fallback != 0 || total + 1 > out_cap;
    if (!bad) for (int j = 0; j < my_total; j++) out[1 + off + j] = scratch[my_first + j];
    __threadfence_system();
    __syncthreads();
    if (tid == 0) { out[0] = bad ? -1 : total; __threadfence_system(); }
}
