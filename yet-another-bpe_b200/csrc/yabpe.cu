// yabpe.cu -- extern "C" entry points of libyabpe.so (see include/yabpe.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <stdio.h>
#include <string.h>

#include "../../include/yabpe.h"
#include "encode.cuh"
#include "pretok_fast.cuh"
#include "decode.cuh"
#include "exchange.cuh"
#include "encode_small.cuh"

static thread_local char g_err[512] = "";
static long long g_launches = 0;

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            snprintf(g_err, sizeof g_err, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return YABPE_ERR_CUDA;                                                              \
        }                                                                                       \
    } while (0)

#define ARG_CHECK(cond)                                                                          \
    do {                                                                                        \
        if (!(cond)) { snprintf(g_err, sizeof g_err, "bad argument: %s (%s:%d)", #cond, __FILE__, __LINE__); return YABPE_ERR_ARG; } \
    } while (0)

#define LAUNCHED() (g_launches++)

// Per-device state.  Function attributes (dynamic shared-memory opt-in), the SM count and occupancy results belong to
// ONE device: a process that switches devices (the Python layer supports it) must not reuse what it learnt on the first.
#define YABPE_MAX_DEVICES 64
struct DevInfo { int num_sms; bool pretok_attr, merge_attr, small_attr; int enc_per_sm[2]; };
static DevInfo g_dev[YABPE_MAX_DEVICES];
static DevInfo& dev_info() {
    int dev = 0; cudaGetDevice(&dev);
    if (dev < 0 || dev >= YABPE_MAX_DEVICES) dev = YABPE_MAX_DEVICES - 1;      // state shared beyond 64 devices: still correct, re-queried lazily
    DevInfo& D = g_dev[dev];
    if (!D.num_sms) {
        cudaDeviceGetAttribute(&D.num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (D.num_sms <= 0) D.num_sms = 148;
    }
    return D;
}
static int num_sms() { return dev_info().num_sms; }

extern "C" const char* yabpe_last_error(void) { return g_err; }
extern "C" int yabpe_abi_version(void) { return YABPE_ABI_VERSION; }
extern "C" int64_t yabpe_launch_count(void) { return g_launches; }
extern "C" int64_t yabpe_sizeof(int32_t which) {
    switch (which) {
        case 0: return (int64_t)sizeof(yabpe_pretok_args);
        case 1: return (int64_t)sizeof(yabpe_word_table);
        case 2: return (int64_t)sizeof(yabpe_merge_args);
        case 3: return (int64_t)sizeof(yabpe_encode_model);
        case 4: return (int64_t)sizeof(yabpe_encode_out);
        case 5: return (int64_t)sizeof(yabpe_decode_args);
        case 6: return (int64_t)sizeof(yabpe_partition_args);
        default: return -1;
    }
}

extern "C" int yabpe_class_of(uint32_t cp) {
    if (cp >= 0x110000) return 0;
    unsigned blk = yabpe_ucd_stage1[cp >> 8];
    unsigned byte = yabpe_ucd_stage2[blk * 64 + ((cp & 255) >> 2)];
    return (byte >> ((cp & 3) * 2)) & 3;
}

extern "C" int yabpe_device_init(void) {
    CUDA_TRY(cudaMemcpyToSymbol(g_ucd_stage1, yabpe_ucd_stage1, sizeof(yabpe_ucd_stage1)));
    CUDA_TRY(cudaMemcpyToSymbol(g_ucd_stage2, yabpe_ucd_stage2, sizeof(yabpe_ucd_stage2)));
    return YABPE_OK;
}

static int upload_specials(const uint8_t* blob, const int32_t* offs, int32_t n, cudaStream_t st) {
    static thread_local SpecialSet h;
    ARG_CHECK(n >= 0 && n <= YABPE_MAX_SPECIALS);
    memset(&h, 0, sizeof h);
    h.n = n;
    for (int i = 0; i <= n; i++) h.offs[i] = n ? offs[i] : 0;
    int total = n ? offs[n] : 0;
    ARG_CHECK(total <= YABPE_MAX_SPECIAL_BYTES);
    if (total) memcpy(h.blob, blob, (size_t)total);
    for (int i = 0; i < n; i++) {
        int len = offs[i + 1] - offs[i];
        ARG_CHECK(len > 0);                    /* empty specials are not supported */
        ARG_CHECK(len <= PT_HR / 2);
        if (len > h.max_len) h.max_len = len;
        unsigned char b = blob[offs[i]];
        if (!((h.first_byte_mask[b >> 3] >> (b & 7)) & 1)) {
            if (h.n_first < 8) h.first[h.n_first] = b;
            h.n_first++;
        }
        h.first_byte_mask[b >> 3] |= (unsigned char)(1u << (b & 7));
    }
    // the same set on the same device and stream as the last upload is already in place (stream order): short encodes
    // would otherwise pay this copy on every call
    static thread_local SpecialSet last; static thread_local int last_dev = -1; static thread_local cudaStream_t last_st = nullptr;
    int dev = 0; cudaGetDevice(&dev);
    if (dev == last_dev && st == last_st && memcmp(&last, &h, sizeof h) == 0) return YABPE_OK;
    CUDA_TRY(cudaMemcpyToSymbolAsync(c_sp, &h, sizeof h, 0, cudaMemcpyHostToDevice, st));
    last = h; last_dev = dev; last_st = st;
    return YABPE_OK;
}

static int make_params(const yabpe_pretok_args* a, PretokParams* P) {
    ARG_CHECK(a && a->text && a->n > 0);
    ARG_CHECK(a->own_lo >= 0 && a->own_hi <= a->n && a->own_lo <= a->own_hi);
    ARG_CHECK(a->short_cap > 0 && (a->short_cap & (a->short_cap - 1)) == 0);
    ARG_CHECK(a->long_cap > 0 && (a->long_cap & (a->long_cap - 1)) == 0);
    ARG_CHECK(a->n_sp == 0 || (a->cand_bits && a->rec_bits));
    ARG_CHECK(((uintptr_t)a->text & 15) == 0);
    P->text = a->text; P->n = a->n; P->cuts = (const i64*)a->cuts; P->n_cuts = a->n_cuts;
    P->mode = a->mode; P->n_sp = a->n_sp;
    P->own_lo = a->own_lo; P->own_hi = a->own_hi;
    P->cand = a->cand_bits; P->rec = a->rec_bits;
    P->scap = a->short_cap;
    P->st.kb = (char*)a->short_keys; P->st.cap = a->short_cap;
    if (a->short_counts) { P->st.ks = 16; P->st.cb = (char*)a->short_counts; P->st.cs = 8; ARG_CHECK(((uintptr_t)a->short_keys & 15) == 0); }
    else { P->st.ks = 32; P->st.cb = (char*)a->short_keys + 16; P->st.cs = 32; ARG_CHECK(((uintptr_t)a->short_keys & 31) == 0); }
    P->lent = (LongEntry*)a->long_entries; P->lcap = a->long_cap;
    P->ovf_pos = (i64*)a->ovf_pos; P->ovf_cap = a->ovf_cap;
    P->stats = (i64*)a->stats;
    P->work = (i64*)a->work; P->work_cap = a->work ? a->work_cap : 0; P->list_mode = 0;
    P->hot_keys = (const uint4*)a->hot_keys;
    P->hot.kb = (char*)a->hot_table; P->hot.cb = (char*)a->hot_table + 16; P->hot.ks = 32; P->hot.cs = 32;
    P->hot.cap = a->hot_table ? a->hot_cap : 0;
    ARG_CHECK(P->hot.cap == 0 || ((P->hot.cap & (P->hot.cap - 1)) == 0 && ((uintptr_t)a->hot_table & 31) == 0));
    P->tile_base = a->own_lo / PT_TILE;
    P->n_tiles = a->own_hi > a->own_lo ? (a->own_hi - 1) / PT_TILE - P->tile_base + 1 : 0;
    return YABPE_OK;
}

extern "C" int32_t yabpe_hot_cache_entries(void) { return PW_NC; }

// hot_keys: PW_NC * 16 bytes, scratch: PW_NC uint64 (both device, zeroed by this call)
extern "C" int yabpe_select_hot(const yabpe_pretok_args* sample, void* hot_keys, void* scratch, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PretokParams P;
    int rc = make_params(sample, &P);
    if (rc) return rc;
    ARG_CHECK(hot_keys && scratch && ((uintptr_t)hot_keys & 15) == 0);
    CUDA_TRY(cudaMemsetAsync(hot_keys, 0, (size_t)PW_NC * 16, st));
    CUDA_TRY(cudaMemsetAsync(scratch, 0, (size_t)PW_NC * 8, st));
    k_hot_select<<<num_sms() * 8, 256, 0, st>>>(P.st, (u64*)scratch, (uint4*)hot_keys, 0); LAUNCHED();
    k_hot_select<<<num_sms() * 8, 256, 0, st>>>(P.st, (u64*)scratch, (uint4*)hot_keys, 1); LAUNCHED();
    CUDA_TRY(cudaGetLastError());
    return YABPE_OK;
}

extern "C" int64_t yabpe_num_tiles(int64_t own_lo, int64_t own_hi) {
    return own_hi > own_lo ? (own_hi - 1) / PT_TILE - own_lo / PT_TILE + 1 : 0;
}

__global__ void __launch_bounds__(256) k_long_tokens_dyn(PretokParams P) {
    // grid-stride over the device-side overflow count
    __shared__ i64 sh_min; __shared__ u64 sh_acc; __shared__ i64 sh[4];
    i64 n_ovf = P.stats[ST_OVF_N];
    if (n_ovf > P.ovf_cap) { if (threadIdx.x == 0 && blockIdx.x == 0) P.stats[ST_TABLE_FULL] = 3; n_ovf = P.ovf_cap; }
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

static int run_specials(const yabpe_pretok_args* a, const PretokParams& P, cudaStream_t st, bool from_own_lo) {
    int rc = upload_specials(a->sp_blob, a->sp_offs, a->n_sp, st);
    if (rc) return rc;
    if (a->n_sp == 0) return YABPE_OK;
    int grid = num_sms() * 8;
    // candidates / resolution need left context for chains that reach into the owned range -- unless the caller declares
    // own_lo a text start (stages bit 4: it is 0 or one of the hard cuts), as the piece-wise upload + count does
    i64 lo = from_own_lo ? P.own_lo : 0, hi = P.n;
    k_special_candidates<<<grid, 256, 0, st>>>(P, lo, hi); LAUNCHED();
    k_resolve_specials<<<grid, 256, 0, st>>>(P, lo, hi, lo); LAUNCHED();
    CUDA_TRY(cudaGetLastError());
    return YABPE_OK;
}

// ---- pre-token starts in text order (the list trainer.py:200-214 returns, as a bitmap) ----
// One thread per byte, the generic start rule over global memory (common.cuh is_token_start: an implementation
// independent of the SWAR scan in k_pretok_warp and of the tile scan in k_pretok_count), one ballot per warp.
__global__ void __launch_bounds__(256) k_token_starts(PretokParams P, uint32_t* bits) {
    GlobalText G{P.text, P.n, P.cuts, P.n_cuts, P.n_sp > 0 ? P.rec : nullptr, -1, P.mode};
    const i64 n_round = (P.n + 31) & ~(i64)31;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < n_round; p += stride) {
        const bool st = p < P.n && is_token_start(G, p);
        const unsigned m = __ballot_sync(0xffffffffu, st);
        if ((threadIdx.x & 31) == 0) bits[p >> 5] = m;
    }
}

extern "C" int yabpe_token_starts(const yabpe_pretok_args* a, uint32_t* start_bits, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PretokParams P;
    int rc = make_params(a, &P);
    if (rc) return rc;
    ARG_CHECK(start_bits != nullptr);
    rc = upload_specials(a->sp_blob, a->sp_offs, a->n_sp, st);
    if (rc) return rc;
    if (P.n <= 0) return YABPE_OK;
    i64 grid = (P.n + 255) / 256;
    if (grid > (i64)num_sms() * 16) grid = (i64)num_sms() * 16;
    k_token_starts<<<(int)grid, 256, 0, st>>>(P, start_bits); LAUNCHED();
    CUDA_TRY(cudaGetLastError());
    return YABPE_OK;
}

extern "C" int yabpe_pretok_count(const yabpe_pretok_args* a, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PretokParams P;
    int rc = make_params(a, &P);
    if (rc) return rc;
    ARG_CHECK(a->ovf_pos && a->ovf_cap > 0 && a->stats);
    int stages = a->stages;
    if ((stages & 7) == 0) stages |= 7;
    if (stages & 1) {
        rc = run_specials(a, P, st, (stages & 16) != 0);
        if (rc) return rc;
    } else {
        rc = upload_specials(a->sp_blob, a->sp_offs, a->n_sp, st);
        if (rc) return rc;
    }
    if (P.n_tiles > 0 && (stages & 2)) {
        DevInfo& DI = dev_info();
        if (!DI.pretok_attr) {
            CUDA_TRY(cudaFuncSetAttribute(k_pretok_count, cudaFuncAttributeMaxDynamicSharedMemorySize, PT_CACHE_BYTES));
            CUDA_TRY(cudaFuncSetAttribute(k_pretok_warp<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PW_SMEM_BYTES));
            CUDA_TRY(cudaFuncSetAttribute(k_pretok_warp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PW_SMEM_BYTES_HOT));
            DI.pretok_attr = true;
        }
        int per_sm = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pretok_count, PT_THREADS, PT_CACHE_BYTES));
        if (per_sm < 1) per_sm = 1;
        // Interior text: the warp-autonomous kernel; it leaves the chunks that touch a hard cut,
        // the ends of the text or the edge of the owned range in P.work for the generic kernel (list mode).
        // Dense cuts (tiny reference chunk sizes) or a missing work list: the generic kernel does everything.
        const i64 c_lo = P.own_lo / PW_CH, c_hi = (P.own_hi - 1) / PW_CH + 1;
        const bool warp_path = !(stages & 8) && P.work && P.work_cap >= 4 * (i64)P.n_cuts + 16 &&
                               c_hi - c_lo >= 4 && 8 * (i64)P.n_cuts < c_hi - c_lo;
        if (warp_path) {
            if (P.hot.cap > 0) {
                i64 grid_w = (c_hi - c_lo + PW_WARPS_HOT - 1) / PW_WARPS_HOT;
                if (grid_w > num_sms()) grid_w = num_sms();
                k_pretok_warp<true><<<(int)grid_w, PW_WARPS_HOT * 32, PW_SMEM_BYTES_HOT, st>>>(P, c_lo, c_hi); LAUNCHED();
                k_hot_flush<<<num_sms() * 4, 256, 0, st>>>(P); LAUNCHED();
            } else {
                i64 grid_w = (c_hi - c_lo + PW_WARPS - 1) / PW_WARPS;
                if (grid_w > num_sms()) grid_w = num_sms();
                k_pretok_warp<false><<<(int)grid_w, PW_THREADS, PW_SMEM_BYTES, st>>>(P, c_lo, c_hi); LAUNCHED();
            }
            PretokParams PL = P;
            PL.list_mode = 1;
            i64 grid = (i64)num_sms() * per_sm;
            if (grid > P.work_cap) grid = P.work_cap;
            k_pretok_count<<<(int)grid, PT_THREADS, PT_CACHE_BYTES, st>>>(PL); LAUNCHED();
        } else {
            int grid = num_sms() * per_sm;
            if ((i64)grid > P.n_tiles) grid = (int)P.n_tiles;
            k_pretok_count<<<grid, PT_THREADS, PT_CACHE_BYTES, st>>>(P); LAUNCHED();
        }
        CUDA_TRY(cudaGetLastError());
    }
    if (P.n_tiles > 0 && (stages & 4)) {
        k_long_tokens_dyn<<<num_sms(), 256, 0, st>>>(P); LAUNCHED();
        CUDA_TRY(cudaGetLastError());
    }
    return YABPE_OK;
}

// ---- merge packed (bytes, offsets, lengths, counts) words into the tables (multi-GPU exchange) ----
__global__ void __launch_bounds__(256) k_insert_words(PretokParams P, const i64* offs, const int32_t* lens, const i64* cnts, i64 nw) {
    u64 my_us = 0, my_ul = 0, my_ub = 0;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 w = (i64)blockIdx.x * blockDim.x + threadIdx.x; w < nw; w += stride) {
        const int len = lens[w];
        const i64 pos = offs[w], c = cnts[w];
        int created = 0;
        if (len <= 0) continue;
        if (len <= PT_SHORT_MAX) {
            u64 lo = 0, hi = 0;
            for (int k = 0; k < len; k++) { u64 b = P.text[pos + k]; if (k < 8) lo |= b << (8 * k); else hi |= b << (8 * (k - 8)); }
            const u64 k0 = (lo & 0x00FFFFFFFFFFFFFFULL) | ((u64)len << 56), k1 = (lo >> 56) | (hi << 8) | (1ULL << 56);
            if (short_insert(P.st, k0, k1, c, &created) < 0) P.stats[ST_TABLE_FULL] = 1;
            if (created) { my_us++; my_ub += len; }
        } else if (len <= 256) {
            u64 h = 0;
            for (int j = 0; j < len; j++) h += long_hash_term(P.text[pos + j], j);
            if (long_insert(P.lent, P.lcap, P.text, long_hash_fix(h), pos, len, c, &created) < 0) P.stats[ST_TABLE_FULL] = 1;
            if (created) { my_ul++; my_ub += len; }
        }
    }
    if (my_us) atomicAdd((u64*)&P.stats[ST_UNIQ_SHORT], my_us);
    if (my_ul) atomicAdd((u64*)&P.stats[ST_UNIQ_LONG], my_ul);
    if (my_ub) atomicAdd((u64*)&P.stats[ST_UNIQ_BYTES], my_ub);
}
__global__ void __launch_bounds__(256) k_insert_words_long(PretokParams P, const i64* offs, const int32_t* lens, const i64* cnts, i64 nw) {
    __shared__ u64 sh_acc; __shared__ i64 sh[4];
    for (i64 w = blockIdx.x; w < nw; w += gridDim.x) {
        const i64 len = lens[w];
        if (len <= 256) continue;
        u64 h = block_long_hash(P.text, offs[w], len, &sh_acc);
        int created;
        i64 slot = block_long_upsert(P.lent, P.lcap, P.text, h, offs[w], len, cnts[w], false, &created, sh);
        if (threadIdx.x == 0) {
            if (slot < 0) P.stats[ST_TABLE_FULL] = 1;
            if (created) { atomicAdd((u64*)&P.stats[ST_UNIQ_LONG], 1ULL); atomicAdd((u64*)&P.stats[ST_UNIQ_BYTES], (u64)len); }
        }
        __syncthreads();
    }
}

extern "C" int yabpe_insert_words(const yabpe_pretok_args* a, const int64_t* offs, const int32_t* lens, const int64_t* counts,
                                  int64_t n_words, int32_t has_long, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PretokParams P;
    int rc = make_params(a, &P);
    if (rc) return rc;
    ARG_CHECK(offs && lens && counts && n_words >= 0 && a->stats);
    if (n_words == 0) return YABPE_OK;
    k_insert_words<<<num_sms() * 8, 256, 0, st>>>(P, (const i64*)offs, lens, (const i64*)counts, n_words); LAUNCHED();
    if (has_long) { k_insert_words_long<<<num_sms() * 2, 256, 0, st>>>(P, (const i64*)offs, lens, (const i64*)counts, n_words); LAUNCHED(); }
    CUDA_TRY(cudaGetLastError());
    return YABPE_OK;
}

extern "C" int yabpe_partition_words(const yabpe_partition_args* p, int32_t pass, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    ARG_CHECK(p && p->n_words >= 0 && p->n_ranks >= 1 && p->n_ranks <= PX_MAX_RANKS);
    if (p->n_words == 0) return YABPE_OK;
    ARG_CHECK(p->words.wsym && p->words.woff && p->words.wlen && p->words.wcnt && p->dest && p->totals);
    ARG_CHECK(p->n_words < (1LL << PX_WORD_BITS));
    PartParams P;
    P.wsym = p->words.wsym; P.woff = (const i64*)p->words.woff; P.wlen = p->words.wlen; P.wcnt = (const i64*)p->words.wcnt;
    P.n_words = p->n_words; P.G = p->n_ranks; P.dest = p->dest; P.totals = (u64*)p->totals;
    P.base_w = (const i64*)p->base_w; P.base_b = (const i64*)p->base_b; P.cursor = (u64*)p->cursor;
    P.out_lens = p->out_lens; P.out_cnts = (i64*)p->out_cnts; P.out_data = p->out_data;
    i64 grid = (p->n_words + 255) / 256;
    if (grid > (i64)num_sms() * 8) grid = (i64)num_sms() * 8;
    if (pass == 0) { k_partition_words<0><<<(int)grid, 256, 0, st>>>(P); LAUNCHED(); }
    else {
        ARG_CHECK(p->base_w && p->base_b && p->cursor && p->out_lens && p->out_cnts && p->out_data);
        k_partition_words<1><<<(int)grid, 256, 0, st>>>(P); LAUNCHED();
    }
    CUDA_TRY(cudaGetLastError());
    return YABPE_OK;
}

static WordTable make_words(const yabpe_word_table* w) {
    WordTable W;
    W.wsym = w->wsym; W.sym_word = w->sym_word; W.woff = (i64*)w->woff; W.wlen = w->wlen; W.wcnt = (i64*)w->wcnt;
    W.sword = w->sword; W.lword = w->lword; W.counters = (i64*)w->counters;
    return W;
}

extern "C" int yabpe_compact_words(const yabpe_pretok_args* a, const yabpe_word_table* w, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    ARG_CHECK(a && w && w->wsym && w->sym_word && w->woff && w->wlen && w->wcnt && w->counters);
    WordTable W = make_words(w);
    int grid = num_sms() * 8;
    PretokParams PC; { int rc0 = make_params(a, &PC); if (rc0) return rc0; }
    k_compact_short<<<grid, 256, 0, st>>>(PC.st, a->short_cap, W); LAUNCHED();
    k_compact_long<<<grid, 256, 0, st>>>((const LongEntry*)a->long_entries, a->long_cap, a->text, W); LAUNCHED();
    CUDA_TRY(cudaGetLastError());
    return YABPE_OK;
}

extern "C" int yabpe_merge_loop(const yabpe_merge_args* m, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    ARG_CHECK(m && m->pcap > 0 && (m->pcap & (m->pcap - 1)) == 0);
    ARG_CHECK(m->tset_cap > 0 && (m->tset_cap & (m->tset_cap - 1)) == 0);
    ARG_CHECK(m->n_syms < 0xffffffffLL && m->pcap < 0x7fffffffLL);
    static_assert(sizeof(Best) == 24, "Best layout");
    MergeParams M;
    M.wsym = m->words.wsym; M.sym_word = m->words.sym_word; M.n_syms = m->n_syms;
    M.woff = (const i64*)m->words.woff; M.wlen = m->words.wlen; M.wcnt = (const i64*)m->words.wcnt; M.n_words = m->n_words; M.wstamp = m->wstamp; M.wslot = m->wslot; M.newp = m->newp;
    M.tok_bytes = m->tok_bytes; M.tok_bytes_cap = m->tok_bytes_cap; M.tok_off = (i64*)m->tok_off;
    M.tok_hash = (u64*)m->tok_hash; M.tok_pow = (u64*)m->tok_pow; M.tok_pre = (u64*)m->tok_pre; M.tset = (u64*)m->tset; M.tset_cap = m->tset_cap;
    M.max_tokens = m->max_tokens;
    M.pkey = (u64*)m->pkey; M.pcnt = (i64*)m->pcnt; M.pcap = m->pcap;
    M.ioff = m->ioff; M.icnt = m->icnt; M.ipost = (i64*)m->ipost; M.inact = m->inact; M.intop = m->intop; M.act = m->act;
    M.top_slot = m->top_slot; M.top_key = (u64*)m->top_key; M.hist = m->hist;
    M.alog_word = (i64*)m->alog_word; M.alog_cap = m->alog_cap; M.seg_start = m->seg_start; M.seg_end = m->seg_end;
    M.merge_next = m->merge_next; M.tok_first = m->tok_first; M.tok_head = (int4*)m->tok_head;
    ARG_CHECK(((uintptr_t)m->tok_head & 15) == 0);
    ARG_CHECK(m->alog_cap >= 2 * m->n_words + ML_LEADER_ITEMS_MAX);
    M.partial = (Best*)m->partial; M.bsum = (i64*)m->bsum;
    M.merges = m->merges; M.merge_new = m->merge_new; M.state = (i64*)m->state;
    M.num_merges = m->num_merges; M.min_freq = m->min_frequency; M.rebuild_every = m->rebuild_every;
    M.helper_mode = m->helper_mode; M.batch_max = m->batch_max;
    M.helper_min_syms = m->helper_min_syms > 0 ? m->helper_min_syms : (m->helper_min_syms < 0 ? 0 : ML_HELPER_MIN_SYMS);

    DevInfo& DI = dev_info();
    if (!DI.merge_attr) {
        CUDA_TRY(cudaFuncSetAttribute(k_merge_loop, cudaFuncAttributeMaxDynamicSharedMemorySize, ML_DYN_SMEM_BYTES));
        DI.merge_attr = true;
    }
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_merge_loop, ML_THREADS, ML_DYN_SMEM_BYTES));
    ARG_CHECK(per_sm >= 1);
    per_sm = 1;
    int grid = num_sms() * per_sm;
    if (grid > 1024) grid = 1024;
    void* args[] = {&M};
    CUDA_TRY(cudaLaunchCooperativeKernel((void*)k_merge_loop, dim3(grid), dim3(ML_THREADS), args, ML_DYN_SMEM_BYTES, st)); LAUNCHED();
    return YABPE_OK;
}

static EncodeModel make_model(const yabpe_encode_model* e) {
    EncodeModel E;
    E.mkey = (const u64*)e->mkey; E.mval = (const u64*)e->mval; E.mcap = e->mcap;
    E.byte_sym = e->byte_sym; E.sym_out = e->sym_out; E.sp_ids = e->sp_ids; E.consistent = e->consistent;
    return E;
}

__global__ void __launch_bounds__(256) k_encode_words_short(EncodeModel E, int32_t* wsym, const i64* woff, int32_t* wlen, i64 n_words) {
    i64 gstride = (i64)gridDim.x * blockDim.x;
    for (i64 w = (i64)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += gstride) {
        int n = wlen[w];
        if (n > ENC_LONG_WORD) continue;
        wlen[w] = encode_word_thread(E, wsym + woff[w], n);
    }
}
// Words longer than ENC_LONG_WORD symbols are rare: every block scans 256 word lengths at a time (coalesced) and only
// enters the block-cooperative encoder for the ones it finds (a serial scan of millions of lengths per block cost
// more than the encoding itself).
__global__ void __launch_bounds__(256) k_encode_words_long(EncodeModel E, int32_t* wsym, int32_t* scratch, const i64* woff,
                                                            const int32_t* wlen_in, int32_t* wlen_out, i64 n_words) {
    __shared__ int sh_i[2 + 8]; __shared__ u64 sh_u;
    __shared__ int sh_long[256]; __shared__ int sh_nlong;
    for (i64 base = (i64)blockIdx.x * 256; base < n_words; base += (i64)gridDim.x * 256) {
        const i64 w = base + threadIdx.x;
        const int n = w < n_words ? wlen_in[w] : 0;
        if (threadIdx.x == 0) sh_nlong = 0;
        __syncthreads();
        if (n > ENC_LONG_WORD) sh_long[atomicAdd(&sh_nlong, 1)] = (int)threadIdx.x;
        __syncthreads();
        const int nl = sh_nlong;
        for (int k = 0; k < nl; k++) {
            const i64 wk = base + sh_long[k];
            int r = encode_word_block(E, wsym + woff[wk], scratch + woff[wk], wlen_in[wk], sh_i, &sh_u);
            __syncthreads();
            if (threadIdx.x == 0) wlen_out[wk] = r;
            __syncthreads();
        }
        __syncthreads();
    }
}

extern "C" int yabpe_encode_words(const yabpe_encode_model* e, const yabpe_word_table* w, int64_t n_words, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    ARG_CHECK(e && w && n_words >= 0);
    if (n_words == 0) return YABPE_OK;
    EncodeModel E = make_model(e);
    // short words first: they only shrink, so the long kernel (original length > ENC_LONG_WORD)
    // afterwards still sees exactly the words the short kernel skipped
    k_encode_words_short<<<num_sms() * 8, 256, 0, st>>>(E, w->wsym, (const i64*)w->woff, w->wlen, n_words); LAUNCHED();
    k_encode_words_long<<<num_sms() * 2, 256, 0, st>>>(E, w->wsym, w->sym_word, (const i64*)w->woff, w->wlen, w->wlen, n_words); LAUNCHED();
    CUDA_TRY(cudaGetLastError());
    return YABPE_OK;
}

extern "C" int yabpe_encode_finalize(const yabpe_pretok_args* a, const yabpe_encode_model* e, const yabpe_word_table* w,
                                     int64_t n_words, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    ARG_CHECK(a && e && w && n_words >= 0 && w->sword && w->lword);
    ARG_CHECK(a->short_cap > 0 && a->long_cap > 0);
    if (n_words == 0) return YABPE_OK;
    EncodeModel E = make_model(e);
    k_encode_finalize_ids<<<num_sms() * 8, 256, 0, st>>>(E, w->wsym, (const i64*)w->woff, w->wlen, n_words); LAUNCHED();
    PretokParams PF; { int rc0 = make_params(a, &PF); if (rc0) return rc0; }
    k_encode_finalize_slots<<<num_sms() * 8, 256, 0, st>>>(PF.st, a->short_cap, w->sword, (LongEntry*)a->long_entries,
                                                          a->long_cap, w->lword, (const i64*)w->woff, w->wlen); LAUNCHED();
    CUDA_TRY(cudaGetLastError());
    return YABPE_OK;
}

extern "C" int yabpe_encode_ids(const yabpe_pretok_args* a, const yabpe_encode_model* e, const yabpe_word_table* w,
                                const yabpe_encode_out* o, int32_t pass, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PretokParams P;
    int rc = make_params(a, &P);
    if (rc) return rc;
    ARG_CHECK(e && w && o && o->tile_count);
    EncodeModel E = make_model(e);
    EncodeOut O;
    O.wsym = w->wsym; O.woff = (const i64*)w->woff; O.wlen = w->wlen; O.sword = w->sword; O.lword = w->lword;
    O.tile_count = (i64*)o->tile_count; O.out_ids = o->out_ids; O.out_cap = o->out_cap; O.doc_off = (i64*)o->doc_off;
    if (P.n_tiles == 0) return YABPE_OK;
    rc = upload_specials(a->sp_blob, a->sp_offs, a->n_sp, st);
    if (rc) return rc;
    // persistent grid: as many CTAs per SM as registers and the ~36 KB tile area allow (5 on sm_100a; the passes are
    // bound by the latency of the table probes, so every resident warp counts)
    int* per_sm = dev_info().enc_per_sm;
    if (per_sm[pass != 0] == 0) {
        int nb = 0;
        CUDA_TRY(pass == 0 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_encode_tiles<false>, PT_THREADS, 0)
                           : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_encode_tiles<true>, PT_THREADS, 0));
        per_sm[pass != 0] = nb < 1 ? 1 : nb;
    }
    int grid = num_sms() * per_sm[pass != 0];
    if ((i64)grid > P.n_tiles) grid = (int)P.n_tiles;
    if (pass == 2) {                         // one pass: counts, chained scan and ids (tile_count: n_tiles + 2 words, zeroed)
        ARG_CHECK(o->out_ids || o->out_cap == 0);
        k_encode_tiles_fused<<<grid, PT_THREADS, 0, st>>>(P, E, O); LAUNCHED();
    } else if (pass == 0) {
        k_encode_tiles<false><<<grid, PT_THREADS, 0, st>>>(P, E, O); LAUNCHED();
        k_scan_tiles<<<1, 1024, 0, st>>>((i64*)o->tile_count, P.n_tiles); LAUNCHED();
    } else {
        ARG_CHECK(o->out_ids || o->out_cap == 0);
        k_encode_tiles<true><<<grid, PT_THREADS, 0, st>>>(P, E, O); LAUNCHED();
    }
    CUDA_TRY(cudaGetLastError());
    return YABPE_OK;
}

// ---- one-launch encode of a short text ----
extern "C" int32_t yabpe_encode_small_max_bytes(void) { return ES_MAX_BYTES; }

extern "C" int yabpe_encode_small(const yabpe_encode_model* e, const uint8_t* text, int32_t n, const uint8_t* sp_blob,
                                  const int32_t* sp_offs, int32_t n_sp, int32_t* scratch, int32_t* out, int32_t out_cap, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    ARG_CHECK(e && text && n > 0 && n <= ES_MAX_BYTES && scratch && out && out_cap >= 1);
    int rc = upload_specials(sp_blob, sp_offs, n_sp, st);
    if (rc) return rc;
    EncodeModel E = make_model(e);
    const int nwords = (n + 31) / 32;
    const size_t smem = (size_t)(((n + 15) & ~15) + 64) + 4 * (size_t)(nwords + 2) * 4;
    DevInfo& DI = dev_info();
    if (!DI.small_attr) {                       // the largest texts need a little more than the default 48 KB
        CUDA_TRY(cudaFuncSetAttribute(k_encode_small, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 << 10));
        DI.small_attr = true;
    }
    k_encode_small<<<1, ES_THREADS, smem, st>>>(E, text, n, n_sp, scratch, out, out_cap); LAUNCHED();
    CUDA_TRY(cudaGetLastError());
    return YABPE_OK;
}

// ---- counters to the host without a copy engine ----
// A few 64-bit words, stored by one warp straight into MAPPED pinned host memory.  A cudaMemcpy of the same words would
// queue behind whatever bulk transfer occupies the device-to-host copy engine (encode_pinned streams GBs of ids while
// the next piece needs its table statistics); a store from an SM does not.
__global__ void k_publish(volatile i64* host_dst, const i64* src, int n_words) {
    for (int i = threadIdx.x; i < n_words; i += 32) host_dst[i] = src[i];
    __threadfence_system();
}

extern "C" int yabpe_publish(void* host_mapped_dst, const void* device_src, int32_t n_words, void* stream) {
    ARG_CHECK(host_mapped_dst && device_src && n_words > 0 && n_words <= 4096);
    k_publish<<<1, 32, 0, (cudaStream_t)stream>>>((volatile i64*)host_mapped_dst, (const i64*)device_src, n_words); LAUNCHED();
    CUDA_TRY(cudaGetLastError());
    return YABPE_OK;
}

// ---- ids as uint16 for the way back to the host (vocabularies of at most 65 536 entries) ----
// The device -> host copy of the ids bounds the host-buffer encode (2 bytes of int32 ids per text byte on English-like text
// against 1 byte of text going up): half the bytes, half the time.  Ids are stored modulo 2^16; the caller checks the range.
__global__ void __launch_bounds__(256) k_narrow_ids(const int32_t* __restrict__ ids, uint16_t* __restrict__ out, i64 n) {
    const i64 stride = (i64)gridDim.x * blockDim.x * 4;
    for (i64 i = ((i64)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 4 <= n && (((uintptr_t)(ids + i)) & 15) == 0 && (((uintptr_t)(out + i)) & 7) == 0) {
            const int4 v = *(const int4*)(ids + i);
            *(uint2*)(out + i) = make_uint2(((uint32_t)v.x & 0xffffu) | ((uint32_t)v.y << 16), ((uint32_t)v.z & 0xffffu) | ((uint32_t)v.w << 16));
        } else {
            for (i64 j = i; j < n && j < i + 4; j++) out[j] = (uint16_t)ids[j];
        }
    }
}

extern "C" int yabpe_narrow_ids(const int32_t* ids, uint16_t* out, int64_t n, void* stream) {
    ARG_CHECK(n >= 0 && (n == 0 || (ids && out)));
    if (n == 0) return YABPE_OK;
    i64 grid = (n / 4 + 255) / 256;
    if (grid > (i64)num_sms() * 16) grid = (i64)num_sms() * 16;
    if (grid < 1) grid = 1;
    k_narrow_ids<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(ids, out, n); LAUNCHED();
    CUDA_TRY(cudaGetLastError());
    return YABPE_OK;
}

// ---- decode: ids -> bytes ----
extern "C" int64_t yabpe_decode_blocks(int64_t n_ids) { return n_ids > 0 ? (n_ids + DC_IDS - 1) / DC_IDS : 0; }

extern "C" int yabpe_decode_ids(const yabpe_decode_args* d, int32_t pass, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    ARG_CHECK(d && d->n_ids >= 0 && d->block_count && d->vocab_cap >= 0);
    if (d->n_ids == 0) return YABPE_OK;
    ARG_CHECK(d->ids && d->tok_off && d->tok_bytes);
    DecodeParams D;
    D.ids = d->ids; D.n_ids = d->n_ids; D.tok_off = (const i64*)d->tok_off; D.tok_bytes = d->tok_bytes;
    D.vocab_cap = d->vocab_cap; D.block_count = (i64*)d->block_count; D.out = d->out; D.out_cap = d->out_cap;
    const i64 n_blocks = yabpe_decode_blocks(d->n_ids);
    i64 grid = (i64)num_sms() * 8;
    if (grid > n_blocks) grid = n_blocks;
    if (pass == 0) {
        k_decode_ids<false><<<(int)grid, DC_THREADS, 0, st>>>(D); LAUNCHED();
        k_scan_tiles<<<1, 1024, 0, st>>>(D.block_count, n_blocks); LAUNCHED();
    } else {
        ARG_CHECK(d->out || d->out_cap == 0);
        k_decode_ids<true><<<(int)grid, DC_THREADS, 0, st>>>(D); LAUNCHED();
    }
    CUDA_TRY(cudaGetLastError());
    return YABPE_OK;
}
