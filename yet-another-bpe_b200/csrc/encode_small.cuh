// encode_small.cuh -- one-launch encode for short inputs (<= ES_MAX_BYTES).
//
// Replaces /root/reference/src/yet_another_bpe/tokenizer.py:152-308 for the strings the reference's own tests pass to
// `encode` (a few bytes to a few KB).  The batched pipeline (encode.cuh) costs 13 launches and three host round trips
// whatever the input length; here ONE CTA does everything and the host waits for one event:
//   text (device or MAPPED pinned host memory) -> shared memory
//   special-token candidates (thread per byte), leftmost / longest-first resolution (tokenizer.py:97-102,171)
//   pre-token start bit per byte: the generic rule of common.cuh (is_token_start), the same code k_token_starts uses
//   thread t owns the pre-tokens that START in bytes [32t, 32t + 32): BPE by rank on each (encode_word_thread, the exact
//     restatement of the heap loop) in thread-local memory, ids compacted into the scratch range the thread's tokens cover
//   block scan of the per-thread id counts -> ids in text order -> out[1..], out[0] = count (or -1: fall back)
// Pre-tokens longer than ES_MAX_TOKEN bytes (the thread-per-token BPE is quadratic) make the kernel report -1; the caller
// then takes the batched path.  No hash tables, no de-duplication: a short text has few repeats.
#pragma once

#include "encode.cuh"

#define ES_MAX_BYTES 32768
#define ES_THREADS 1024
#define ES_MAX_TOKEN 64
#define ES_WORDS (ES_MAX_BYTES / 32)             // one 32-byte segment per thread

static_assert(ES_WORDS == ES_THREADS, "thread t owns the pre-tokens starting in segment t");

__device__ __forceinline__ int es_next_start(const uint32_t* sbits, int from, int n) {      // first start > from, or n
    int i = from + 1;
    while (i < n) {
        const uint32_t w = sbits[i >> 5] >> (i & 31);
        if (w) { const int r = i + __ffs(w) - 1; return r < n ? r : n; }
        i = ((i >> 5) + 1) << 5;
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
    __shared__ int sh_scan[ES_THREADS / 32];
    __shared__ int sh_fallback;
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

    // ---- thread t: the pre-tokens starting in [32t, 32t + 32)
    int my_total = 0, my_first = -1;
    if (tid < nwords) {
        uint32_t m = sbits[tid];
        int cursor = -1;
        while (m) {
            const int s = (tid << 5) + __ffs(m) - 1; m &= m - 1;
            if (cursor < 0) { cursor = s; my_first = s; }
            if (n_sp > 0 && ((rec[s >> 5] >> (s & 31)) & 1u)) {          // a special: its id, or dropped (tokenizer.py:177-181)
                const int sp = special_match(stxt, s, n);
                const int32_t id = sp >= 0 ? E.sp_ids[sp] : -1;
                if (id >= 0) scratch[cursor++] = id;
                continue;
            }
            const int e = es_next_start(sbits, s, n), len = e - s;
            if (len > ES_MAX_TOKEN) { sh_fallback = 1; continue; }
            int32_t sym[ES_MAX_TOKEN];
            for (int j = 0; j < len; j++) sym[j] = stxt[s + j];
            const int cnt = encode_word_thread(E, sym, len);
            for (int j = 0; j < cnt; j++) scratch[cursor++] = E.sym_out[sym[j]];
        }
        if (my_first >= 0) my_total = cursor - my_first;
    }
    // ---- ids in text order
    int inc = my_total;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
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
