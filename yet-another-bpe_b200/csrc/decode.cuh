// Token ids -> bytes (tokenizer.py:323-349: ids that are not in the vocabulary are skipped, the byte strings
// of the others are concatenated).  The strict / replacing UTF-8 decode of the result stays on the host.
//
// Two passes over the ids, like the encoder's output side: pass 0 sums the byte lengths of DC_IDS ids per CTA,
// k_scan_tiles turns the sums into output offsets, pass 1 recomputes the lengths, scans them inside the CTA and
// writes.  The write loop runs over OUTPUT bytes (4 consecutive bytes per thread, neighbouring threads neighbouring
// bytes), not over tokens: a binary search in the CTA's offset array finds the token a byte belongs to, so the stores
// coalesce whatever the token lengths are; the vocabulary pool (a few hundred KB) is served by L1 / L2.
#pragma once
#include "common.cuh"

#define DC_THREADS 256
#define DC_IPT 8
#define DC_IDS (DC_THREADS * DC_IPT)

struct DecodeParams {
    const int32_t* ids; i64 n_ids;
    const i64* tok_off;          // vocab_cap + 1 offsets into tok_bytes; an id without bytes has an empty range
    const uint8_t* tok_bytes;
    int32_t vocab_cap;
    i64* block_count;            // n_blocks + 1
    uint8_t* out; i64 out_cap;
};

template <bool WRITE>
__global__ void __launch_bounds__(DC_THREADS) k_decode_ids(DecodeParams D) {
    __shared__ uint32_t sh_off[DC_IDS + 1];      // byte offset of every id of this CTA inside the CTA's output
    __shared__ uint32_t sh_warp[DC_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const i64 n_blocks = (D.n_ids + DC_IDS - 1) / DC_IDS;
    for (i64 blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        const i64 id0 = blk * DC_IDS + (i64)tid * DC_IPT;
        uint32_t len[DC_IPT], sum = 0;
#pragma unroll
        for (int k = 0; k < DC_IPT; k++) {
            len[k] = 0;
            if (id0 + k < D.n_ids) {
                const int32_t id = D.ids[id0 + k];
                if (id >= 0 && id < D.vocab_cap) len[k] = (uint32_t)(D.tok_off[id + 1] - D.tok_off[id]);
            }
            sum += len[k];
        }
        uint32_t inc = sum;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) sh_warp[wid] = inc;
        __syncthreads();
        uint32_t wbase = 0, total = 0;
#pragma unroll
        for (int k = 0; k < DC_THREADS / 32; k++) { const uint32_t t = sh_warp[k]; if (k < wid) wbase += t; total += t; }
        if (!WRITE) {
            if (tid == 0) D.block_count[blk] = total;
        } else {
            uint32_t o = wbase + inc - sum;
#pragma unroll
            for (int k = 0; k < DC_IPT; k++) { sh_off[tid * DC_IPT + k] = o; o += len[k]; }
            if (tid == DC_THREADS - 1) sh_off[DC_IDS] = o;
            __syncthreads();
            const i64 obase = D.block_count[blk];
            for (uint32_t j0 = (uint32_t)tid * 4; j0 < total; j0 += DC_THREADS * 4) {
                int lo = 0, hi = DC_IDS;                      // last id with sh_off[id] <= j0 (empty ids share an offset)
                while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (sh_off[mid] <= j0) lo = mid; else hi = mid; }
                int t = lo;
                i64 src = 0; uint32_t t_end = 0; bool have = false;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t j = j0 + q;
                    if (j >= total) break;
                    if (!have || j >= t_end) {
                        have = true;
                        while (sh_off[t + 1] <= j) t++;
                        t_end = sh_off[t + 1];
                        src = D.tok_off[D.ids[blk * DC_IDS + t]] - (i64)sh_off[t];
                    }
                    if (obase + j < D.out_cap) D.out[obase + j] = D.tok_bytes[src + j];
                }
            }
        }
        __syncthreads();
    }
}
