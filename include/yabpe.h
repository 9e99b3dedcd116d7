/*
 * yabpe.h -- C ABI of libyabpe.so: the B200 (sm_100a) implementation of the BPE hot path of
 * DreamOneX/yet-another-bpe.
 *
 * The reference has no FFI of its own (pure Python); the boundary it exposes for this path is
 * tests/adapters.py:37-99 (run_train_bpe / get_tokenizer).  Every entry point below names the
 * reference code it replaces; the Python host layer (yet-another-bpe_b200/yabpe/) mirrors the
 * reference's classes on top of these calls, and INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - every function returns 0 on success, a negative YABPE_ERR_* otherwise;
 *     yabpe_last_error() returns a static, thread-local description of the last failure
 *   - all data pointers are CALLER-OWNED DEVICE pointers unless the comment says "host"
 *   - `stream` is a cudaStream_t passed as void*; all launches are stream-ordered and
 *     non-blocking; nothing synchronises with the host
 *   - buffers whose size is data dependent follow count -> allocate -> fill: the counting
 *     call leaves sizes in a device `stats` / `counters` array the host reads once
 *   - one host thread per device; no global state besides the per-device Unicode tables and
 *     the special-token set in constant memory (set per call, stream-ordered)
 */
#ifndef YABPE_H
#define YABPE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YABPE_ABI_VERSION 6

#define YABPE_OK 0
#define YABPE_ERR_CUDA (-1)
#define YABPE_ERR_ARG (-2)

/* indices into the int64[16] pre-token statistics array */
#define YABPE_ST_NTOK 0        /* pre-token occurrences counted                           */
#define YABPE_ST_UNIQ_SHORT 1  /* unique pre-tokens of <= 14 bytes                        */
#define YABPE_ST_UNIQ_LONG 2   /* unique pre-tokens of  > 14 bytes                        */
#define YABPE_ST_UNIQ_BYTES 3  /* sum of the byte lengths of all unique pre-tokens        */
#define YABPE_ST_ERR_POS 4     /* first invalid UTF-8 offset (INT64_MAX = valid)          */
#define YABPE_ST_TABLE_FULL 5  /* != 0: a hash table overflowed, retry with larger tables */
#define YABPE_ST_OVF_N 6       /* pre-tokens longer than the tile window                  */
#define YABPE_ST_NSPECIAL 8    /* recognised special-token occurrences                    */
#define YABPE_ST_SLOW_N 9      /* boundary work items the warp kernel left to the generic kernel */
#define YABPE_ST_CACHE_HIT 10  /* pre-tokens counted in shared memory by the warp kernel   */

/* indices into the int64[64] merge-loop state array */
#define YABPE_MS_NMERGES 0
#define YABPE_MS_NTOK 1
#define YABPE_MS_ERROR 2       /* bit 0: pair table full, bit 1: token pool full, bit 2: internal */
#define YABPE_MS_NPAIRS 6
#define YABPE_MS_POOL_USED 8
#define YABPE_MS_REBUILDS 9
#define YABPE_MS_TREBUILDS 10
#define YABPE_MS_LEADER_MERGES 15 /* merges run by the single-CTA fast path */
#define YABPE_MS_GRID_MERGES 16
/* statistics (informational): [40..47] phase clocks of CTA 0 in cycles (initial histogram, first index + active set, top-list
 * rebuilds, index rebuilds, leader sessions, grid-mode merges, total, number of top-list rebuilds), [48..53] grid-mode single
 * merges by candidate count (merges, cycles), [54] leader iterations, [55] merges the leader did in batches of two or more,
 * [56] grid-mode batches, [57] merges in them, [58..62] cycles of the grid-mode batches (selection, barrier 1, ranges + lookups,
 * CTA 0's rewrite share, barrier 2) */
#define YABPE_MS_LEADER_ITERS 54
#define YABPE_MS_LEADER_BATCHED 55
#define YABPE_MS_GRID_BATCHES 56
#define YABPE_MS_GRID_BATCHED 57

const char* yabpe_last_error(void);
int yabpe_abi_version(void);

/* Upload the Unicode class tables (regex-module \p{L} \p{N} \s, see tools/gen_unicode_tables.py)
 * to the current device.  Idempotent.  Replaces the tables compiled into the `regex` C
 * extension the reference calls at trainer.py:169 / tokenizer.py:95. */
int yabpe_device_init(void);

/* Class code of one code point from the built-in table (host-side, for tests):
 * 0 = other, 1 = \p{L}, 2 = \p{N}, 3 = \s. */
int yabpe_class_of(uint32_t cp);

/* ---------------------------------------------------------------------------------------------
 * Pre-tokenise + count.   Replaces trainer.py:146-170 (process_chunk: strict UTF-8 decode,
 * regex.findall of "sp1|..|GPT2_PAT") + trainer.py:221-225 (word_freq), and for mode 1 the
 * special split + findall of tokenizer.py:169-186.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    const uint8_t* text;        /* device; readable capacity >= round_up(n, 16) + 16             */
    int64_t n;                  /* bytes of text                                                  */
    const int64_t* cuts;        /* device, sorted, strictly inside (0, n): hard text boundaries   */
    int32_t n_cuts;             /*   = reference chunk ends (trainer.py:172-198), file / doc ends */
    int32_t mode;               /* 0 = trainer (specials are leading alternatives), 1 = encode    */
    const uint8_t* sp_blob;     /* HOST: special tokens, priority order, concatenated             */
    const int32_t* sp_offs;     /* HOST: n_sp + 1 offsets into sp_blob                            */
    int32_t n_sp;
    int32_t stages;             /* 0 = all; else bit 0: special resolution, bit 1: tile kernel,   */
                                /*   bit 2: over-long pre-tokens (lets a caller time each stage), */
                                /*   bit 3: generic tile kernel only (no warp kernel; for tests), */
                                /*   bit 4: own_lo is a text start (0 or one of `cuts`): the special-token  */
                                /*   passes cover [own_lo, n) only -- piece-wise counting into ONE table set */
                                /*   while the text is still being uploaded (caller zeroes stats[6], [9])   */
    int64_t own_lo, own_hi;     /* only pre-tokens starting in [own_lo, own_hi) are counted       */
    uint32_t* cand_bits;        /* device, (n + 63) / 32 words, zeroed; may be NULL when n_sp = 0 */
    uint32_t* rec_bits;         /* device, same size, zeroed: recognised special starts (output)  */
    void* short_keys;           /* device, zeroed.  short_counts != NULL: short_cap * 16 bytes of keys;      */
                                /*   short_counts == NULL: 32-byte aligned, short_cap * 32 bytes, per slot      */
                                /*   {key word 0, key word 1, count, reserved} (key and count in one sector:    */
                                /*   for tables far larger than the L2)                                        */
    int64_t* short_counts;      /* device, short_cap, zeroed -- or NULL for the interleaved layout */
    int64_t short_cap;          /* power of two                                                   */
    void* long_entries;         /* device, long_cap * 32 bytes, zeroed                            */
    int64_t long_cap;           /* power of two                                                   */
    int64_t* ovf_pos;           /* device scratch for over-long pre-token starts                  */
    int64_t ovf_cap;
    int64_t* stats;             /* device int64[16]; caller zeroes it and sets [ERR_POS]=INT64_MAX */
    const void* hot_keys;       /* device, yabpe_hot_cache_entries() * 16 bytes from yabpe_select_hot, or NULL: */
                                /*   keys the warp kernel pre-loads into its shared-memory cache                 */
    int64_t* work;              /* device scratch, 3 * work_cap int64 (boundary work items), or NULL */
    int64_t work_cap;           /*   >= 4 * n_cuts + 16 enables the warp kernel in trainer mode   */
    void* hot_table;            /* device, 32-byte aligned, hot_cap * 32 bytes, zeroed once -- or NULL.  A direct-mapped table small   */
    int64_t hot_cap;            /*   enough to stay in the L2 (power of two, e.g. 2^20 slots = 32 MB), probed before the big short table */
                                /*   by the warp kernel when the big one is DRAM-sized; its counts are folded into the big table at   */
                                /*   the end of every call, its claimed keys serve later calls on the same tables.  Result-neutral   */
} yabpe_pretok_args;

/* Stage 1: special-token candidates + resolution (skipped when n_sp == 0).
 * Stage 2: tile kernel: classes, token starts, keys, hash-table counts.
 * Stage 3 (yabpe_pretok_finish): pre-tokens longer than the tile window; needs stats[OVF_N],
 *          which it reads on the device -- no host round trip. */
int yabpe_pretok_count(const yabpe_pretok_args* a, void* stream);

/* Pre-token starts in text order: bit p of start_bits (device, (n + 31) / 32 words) is set iff a pre-token starts
 * at byte p.  Replaces the ORDER of the list trainer.py:200-214 (_preprocess_corpus) returns -- the counting path
 * only keeps the multiset.  Call after yabpe_pretok_count on the same args (it reads rec_bits, the recognised
 * specials, and relies on that call's strict UTF-8 check). */
int yabpe_token_starts(const yabpe_pretok_args* a, uint32_t* start_bits, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Word table.  Replaces the dict[tuple[bytes,...], int] built at trainer.py:221-225 by flat
 * arrays: one int32 symbol slot per byte of every unique pre-token.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t* wsym;              /* device, n_syms slots: current symbols of every word            */
    int32_t* sym_word;          /* device, n_syms slots: owning word of each slot                 */
    int64_t* woff;              /* device, n_words: first slot of each word                       */
    int32_t* wlen;              /* device, n_words: current length                                */
    int64_t* wcnt;              /* device, n_words: occurrences                                   */
    int32_t* sword;             /* device, short_cap: slot -> word id (may be NULL)               */
    int32_t* lword;             /* device, long_cap : slot -> word id (may be NULL)               */
    int64_t* counters;          /* device int64[2], zeroed: [0] n_words, [1] n_syms (outputs)     */
} yabpe_word_table;

int yabpe_compact_words(const yabpe_pretok_args* a, const yabpe_word_table* w, void* stream);

/* Multi-GPU exchange: add `n_words` packed pre-tokens (bytes a->text[offs[i] .. offs[i]+lens[i]), occurrence
 * counts[i]) to the tables in `a` (a->text is the packed byte blob; long entries reference it).  Used
 * after the NCCL all-to-all of hash-partitioned (word, count) lists; merges duplicates from different
 * ranks exactly like trainer.py:221-225 merges occurrences.  has_long != 0 when some lens[i] > 256. */
int yabpe_insert_words(const yabpe_pretok_args* a, const int64_t* offs, const int32_t* lens, const int64_t* counts,
                       int64_t n_words, int32_t has_long, void* stream);

/* Multi-GPU exchange, sending side: partition the unique words of `w` (straight after yabpe_compact_words, symbols are
 * still single bytes) by hash(bytes) mod n_ranks and pack them per destination rank for the NCCL all-to-all.
 *   pass 0: dest[i] = destination of word i; totals[d] = (bytes << 26 | words) sent to rank d (totals zeroed by the caller)
 *   pass 1: out_lens / out_cnts / out_data filled, destination d's words contiguous from word base_w[d] / byte base_b[d]
 *           (base_* = exclusive prefix sums of the unpacked totals, device); cursor = n_ranks zeroed uint64 scratch
 * Equal byte strings get equal destinations on every rank, so each rank can merge its partition on its own
 * (yabpe_insert_words) -- together they rebuild trainer.py:221-225's word_freq of the whole corpus. */
typedef struct {
    yabpe_word_table words; int64_t n_words;
    int32_t n_ranks; int32_t _pad;
    uint8_t* dest;              /* device, n_words                                                 */
    uint64_t* totals;           /* device, n_ranks                                                 */
    const int64_t* base_w; const int64_t* base_b;   /* device, n_ranks each (pass 1)               */
    uint64_t* cursor;           /* device, n_ranks, zeroed (pass 1)                                */
    int32_t* out_lens; int64_t* out_cnts; uint8_t* out_data;   /* device (pass 1)                  */
} yabpe_partition_args;

int yabpe_partition_words(const yabpe_partition_args* p, int32_t pass, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Merge loop.  Replaces trainer.py:216-302 (_merge_loop): pair histogram, best-pair selection
 * max(count, (left_bytes, right_bytes)), left-to-right rewrite, incremental deltas.  One
 * cooperative launch runs all merges; the host reads `state`, `merges` and the token pool after.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    yabpe_word_table words; int64_t n_words, n_syms;
    int32_t* wstamp;            /* device, n_words, zeroed                                        */
    int32_t* wslot;             /* device, n_syms + 8: cached pair-table slot of every adjacency  */
    int32_t* newp;              /* device, n_syms + 8: scratch (pairs created by one merge)        */
    uint8_t* tok_bytes; int64_t tok_bytes_cap;   /* token byte pool; base tokens filled by host   */
    int64_t* tok_off;           /* device, max_tokens + 1; [0..n_base] filled by the host         */
    uint64_t* tok_hash;         /* device, max_tokens; polynomial hash (base 0x100000001b3, +1)   */
    uint64_t* tok_pow;          /* device, max_tokens; base^len                                   */
    uint64_t* tok_pre;          /* device, max_tokens; scratch (8-byte big-endian prefix per token) */
    uint64_t* tset; int64_t tset_cap;            /* device hash set of token hashes, pow2         */
    int64_t max_tokens;
    uint64_t* pkey; int64_t* pcnt; int64_t pcap; /* pair table, pow2, zeroed                      */
    uint32_t* ioff;             /* device, pcap + 1                                               */
    uint32_t* icnt;             /* device, pcap                                                   */
    int64_t* ipost;             /* device, n_syms + 8: postings (word | first symbol slot << 32)   */
    uint32_t* inact;            /* device, (pcap + 31) / 32                                       */
    uint32_t* intop;            /* device, (pcap + 31) / 32, zeroed                               */
    int32_t* top_slot; uint64_t* top_key;        /* device, 1024 entries each                     */
    int32_t* hist;              /* device, 1024 ints                                              */
    int32_t* act;               /* device, pcap                                                   */
    int64_t* alog_word; int64_t alog_cap;        /* affected-word log (same entries), >= 2 * n_words + 4096 */
    int32_t* seg_start; int32_t* seg_end;        /* device, num_merges each                       */
    int32_t* merge_next;        /* device, num_merges                                             */
    int32_t* tok_first;         /* device, max_tokens                                             */
    void* tok_head;             /* device, max_tokens * 16 bytes (16-byte aligned scratch)        */
    void* partial;              /* device, 24 bytes per CTA (>= 1024 entries)                     */
    int64_t* bsum;              /* device, one per CTA (>= 1024 entries)                          */
    int32_t* merges;            /* device, 2 * num_merges: (left id, right id) per merge          */
    int32_t* merge_new;         /* device, num_merges: resulting id                               */
    int64_t* state;             /* device int64[64], zeroed except [MS_NTOK] = n_base             */
    int64_t num_merges; int64_t min_frequency;
    int64_t rebuild_every;      /* merges between pair->words index rebuilds; 0 = only when the log is full.
                                 * Candidates of a pair are its postings at the last rebuild plus every word rewritten
                                 * when one of its tokens was created: a superset that grows stale between rebuilds */
    int64_t helper_min_syms;    /* leader mode: idle CTAs prefetch the next merges' words into the L2 when n_syms exceeds this
                                 * (0 = default, 8 Mi slots: smaller word arrays stay L2-resident anyway; < 0 = always).  Result-neutral */
    int64_t helper_mode;        /* 0 = no helpers, 1 = CTAs 1..2, 2 = the CTAs on the SMs next to the leader's                      */
    int64_t batch_max;          /* merges taken per iteration when they provably do not interact (csrc/merge.cuh, "batched merges"):
                                 * 0 = default (31 in grid mode, 16 in leader mode), 1 = strictly one by one as trainer.py:241-300.  Result-neutral                   */
} yabpe_merge_args;

int yabpe_merge_loop(const yabpe_merge_args* m, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Encode.  Replaces tokenizer.py:152-308.  `yabpe_encode_words` applies the merges by rank to
 * every unique pre-token of the batch (in place in wsym); `yabpe_encode_ids` makes the two
 * passes over the text that count and then write the ids in text order.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    const uint64_t* mkey; const uint64_t* mval; int64_t mcap;  /* (sym,sym) -> rank<<32 | result  */
    const int32_t* byte_sym;    /* device, 256                                                    */
    const int32_t* sym_out;     /* device, n_symbols: vocab id (unk substituted)                  */
    const int32_t* sp_ids;      /* device, n_sp: vocab id of each special or -1                   */
    int32_t consistent; int32_t _pad;
} yabpe_encode_model;

int yabpe_encode_words(const yabpe_encode_model* e, const yabpe_word_table* w, int64_t n_words, void* stream);

/* Between yabpe_encode_words and yabpe_encode_ids: turns the encoded symbols into vocabulary ids in place and
 * overwrites the occurrence count of every table slot (short table and long_entries[].count) with the lookup
 * record of its word (first id slot << 24 | number of ids).  Needs w->sword / w->lword.  Call exactly once. */
int yabpe_encode_finalize(const yabpe_pretok_args* a, const yabpe_encode_model* e, const yabpe_word_table* w,
                          int64_t n_words, void* stream);

typedef struct {
    int64_t* tile_count;        /* device, n_tiles + 1                                            */
    int32_t* out_ids; int64_t out_cap;
    int64_t* doc_off;           /* device, n_cuts + 2 (or NULL)                                   */
} yabpe_encode_out;

/* pass 0: per-tile id counts + exclusive scan (total in tile_count[n_tiles]);
 * pass 1: write ids (out_ids must hold tile_count[n_tiles] entries);
 * pass 2: both in ONE pass (decoupled look-back): tile_count has n_tiles + 2 words, ZEROED; out_ids is sized by a guess,
 *         ids beyond out_cap are dropped and tile_count[n_tiles] (the total) tells the caller to repeat with that size. */
int yabpe_encode_ids(const yabpe_pretok_args* a, const yabpe_encode_model* e, const yabpe_word_table* w,
                     const yabpe_encode_out* o, int32_t pass, void* stream);
int64_t yabpe_num_tiles(int64_t own_lo, int64_t own_hi);

/* ---------------------------------------------------------------------------------------------
 * Decode.  Replaces the byte gather of tokenizer.py:323-349 (b"".join(vocab_inv[i] for i in ids if i in vocab_inv));
 * the UTF-8 decode of the result (strict, else errors="replace") stays with the caller.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    const int32_t* ids; int64_t n_ids;   /* device                                                          */
    const int64_t* tok_off;     /* device, vocab_cap + 1: bytes of id i = tok_bytes[tok_off[i] .. tok_off[i+1]);     */
                                /*   an id that is not in the vocabulary has an empty range                          */
    const uint8_t* tok_bytes;   /* device                                                                            */
    int32_t vocab_cap; int32_t _pad;     /* ids outside [0, vocab_cap) are skipped                                   */
    int64_t* block_count;       /* device, yabpe_decode_blocks(n_ids) + 1                                            */
    uint8_t* out; int64_t out_cap;       /* device (pass 1)                                                          */
} yabpe_decode_args;

/* pass 0: byte counts per block + exclusive scan (total bytes in block_count[yabpe_decode_blocks(n_ids)]);
 * pass 1: write the bytes (out must hold that total). */
int yabpe_decode_ids(const yabpe_decode_args* d, int32_t pass, void* stream);
int64_t yabpe_decode_blocks(int64_t n_ids);

/* Short texts (n <= yabpe_encode_small_max_bytes()): the whole of tokenizer.py:152-308 in ONE launch of one CTA -- special
 * split, pre-tokenisation (the generic start rule), BPE by rank per pre-token, ids in text order.  `text` and `out` may be
 * device memory or MAPPED pinned host memory (then the call needs no copy at all); scratch: n int32 (device).
 * out[0] = number of ids, out[1..] = the ids; out[0] = -1 when a pre-token is longer than the kernel handles (64 bytes)
 * or out_cap (>= n + 1 is always enough) is too small: the caller then uses the batched path.  sp_* as in yabpe_pretok_args,
 * already sorted longest-first (tokenizer.py:99). */
int yabpe_encode_small(const yabpe_encode_model* e, const uint8_t* text, int32_t n, const uint8_t* sp_blob,
                       const int32_t* sp_offs, int32_t n_sp, int32_t* scratch, int32_t* out, int32_t out_cap, void* stream);
int32_t yabpe_encode_small_max_bytes(void);

/* ids (int32, device) narrowed to uint16 (device) for the copy back to the host: for vocabularies of at most 65 536 entries
 * the id download that bounds a host-buffer encode halves.  No reference counterpart (tokenizer.py returns list[int]). */
int yabpe_narrow_ids(const int32_t* ids, uint16_t* out, int64_t n, void* stream);

/* Hot set for the warp kernel's shared-memory cache: from the tables of a counted SAMPLE of the corpus (`sample`, after
 * yabpe_pretok_count on e.g. its first 16 MB) pick, for every cache index, the most frequent key that maps to it.
 * hot_keys: yabpe_hot_cache_entries() * 16 bytes; scratch: yabpe_hot_cache_entries() * 8 bytes.  Result-neutral. */
int32_t yabpe_hot_cache_entries(void);
int yabpe_select_hot(const yabpe_pretok_args* sample, void* hot_keys, void* scratch, void* stream);

/* Store n_words (<= 4096) int64 device words into MAPPED pinned host memory (cudaHostAlloc / torch pin_memory under
 * unified addressing) from a kernel: stream-ordered like a copy, but it does not wait for the copy engines, which
 * BBPETokenizer.encode_pinned keeps busy with bulk transfers.  The host reads them after an event on `stream`.
 * (Replaces the `.item()` / `.cpu()` reads of table statistics; no reference counterpart.) */
int yabpe_publish(void* host_mapped_dst, const void* device_src, int32_t n_words, void* stream);

/* kernels launched by this library since load (the bench's `gpu_launches`) */
int64_t yabpe_launch_count(void);

/* sizeof() of the argument structs as compiled (0 pretok_args, 1 word_table, 2 merge_args, 3 encode_model,
 * 4 encode_out, 5 decode_args, 6 partition_args): lets a binding check its own struct layout against the library it loaded. */
int64_t yabpe_sizeof(int32_t which);

#ifdef __cplusplus
}
#endif
#endif
