// exchange.cuh -- packing of the unique (word, count) list for the multi-GPU exchange.
//
// After the local count every rank holds its unique pre-tokens as flat word arrays (k_compact_*).  Before the NCCL
// all-to-all they are partitioned by hash(word bytes) mod G and packed per destination as
//   lens[int32]  cnts[int64]  data[uint8]     (words of destination d contiguous, same order in all three)
// pass 0  per word: destination (stored in `dest`), per-destination word and byte totals
// pass 1  per word: slot and byte offset inside its destination's segment from ONE packed 64-bit cursor per
//         destination (bytes << 26 | words: both offsets of a warp's group are claimed by a single atomic, so the
//         lens order and the data order agree), then the bytes are narrowed int32 -> uint8 into place
// The reference has no counterpart (it is single-process); the merged table equals trainer.py:221-225's word_freq.
#pragma once

#include "common.cuh"

#define PX_MAX_RANKS 64
#define PX_WORD_BITS 26            // words per destination < 2^26, bytes per destination < 2^38

struct PartParams {
    const int32_t* wsym; const i64* woff; const int32_t* wlen; const i64* wcnt; i64 n_words;
    int G;
    uint8_t* dest;                 // n_words: destination rank of every word (pass 0 -> pass 1)
    u64* totals;                   // G packed totals (pass 0 output)
    const i64* base_w; const i64* base_b;    // G exclusive prefix sums of the totals (pass 1 input)
    u64* cursor;                   // G packed cursors, zeroed (pass 1)
    int32_t* out_lens; i64* out_cnts; uint8_t* out_data;
};

// FNV-1a over the bytes, finalised: a function of the BYTES only, so equal words meet on one rank
__device__ __forceinline__ u64 word_hash_syms(const int32_t* s, int n) {
    u64 h = 0xcbf29ce484222325ULL;
    for (int k = 0; k < n; k++) { h ^= (u64)(uint32_t)s[k] & 0xffu; h *= 0x100000001b3ULL; }
    return mix64(h);
}

template <int PASS>
__global__ void __launch_bounds__(256) k_partition_words(PartParams P) {
    __shared__ u64 sh_tot[PX_MAX_RANKS];
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    if (PASS == 0) { if (threadIdx.x < PX_MAX_RANKS) sh_tot[threadIdx.x] = 0; __syncthreads(); }
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 base = (i64)blockIdx.x * blockDim.x; base < P.n_words; base += stride) {       // warp-uniform trip count
        const i64 w = base + threadIdx.x;
        const bool ok = w < P.n_words;
        const int len = ok ? P.wlen[w] : 0;
        const i64 off = ok ? P.woff[w] : 0;
        int d = -1;
        if (ok) d = PASS == 0 ? (int)(word_hash_syms(P.wsym + off, len) % (u64)P.G) : (int)P.dest[w];
        if (PASS == 0) {
            if (ok) { P.dest[w] = (uint8_t)d; atomicAdd(&sh_tot[d], ((u64)len << PX_WORD_BITS) | 1ULL); }
            continue;
        }
        // pass 1: lanes with the same destination claim their slots together
        i64 slot = 0, boff = 0;
        uint32_t todo = __ballot_sync(0xffffffffu, ok);
        while (todo) {
            const int leader = __ffs(todo) - 1;
            const int dd = __shfl_sync(0xffffffffu, d, leader);
            const uint32_t grp = __ballot_sync(0xffffffffu, ok && d == dd);
            int inc = (ok && d == dd) ? len : 0;                 // inclusive prefix of the group's byte lengths
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            const int tot_b = __shfl_sync(0xffffffffu, inc, 31);
            u64 got = 0;
            if (lane == leader) got = atomicAdd(&P.cursor[dd], ((u64)tot_b << PX_WORD_BITS) | (u64)__popc(grp));
            got = __shfl_sync(0xffffffffu, got, leader);
            if (ok && d == dd) {
                slot = P.base_w[dd] + (i64)(got & ((1ULL << PX_WORD_BITS) - 1)) + __popc(grp & lt);
                boff = P.base_b[dd] + (i64)(got >> PX_WORD_BITS) + inc - len;
            }
            todo &= ~grp;
        }
        if (ok) {
            P.out_lens[slot] = len; P.out_cnts[slot] = P.wcnt[w];
            const int32_t* s = P.wsym + off;
            for (int k = 0; k < len; k++) P.out_data[boff + k] = (uint8_t)s[k];
        }
    }
    if (PASS == 0) {
        __syncthreads();
        if (threadIdx.x < P.G && sh_tot[threadIdx.x]) atomicAdd(&P.totals[threadIdx.x], sh_tot[threadIdx.x]);
    }
}
