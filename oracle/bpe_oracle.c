/*
 * bpe_oracle.c -- CPU restatement of DreamOneX/yet-another-bpe's hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA
 * implementation under yet-another-bpe_b200/.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * path never links, imports or calls anything in oracle/.
 *
 * Parity is PINNED (see oracle/README.md and tests/test_oracle_golden.py):
 *   - tests/fixtures_gpt2/train-bpe-reference-merges.txt (243 merges, corpus.en @ 500)
 *   - tests/_snapshots/test_train_bpe_special_tokens.pkl  (structure; input blob missing upstream)
 *   - tests/golden/ (vectors produced by running the reference itself,
 *     generator: tools/make_golden.py)
 *   - in the authoring container: direct comparison with the imported reference
 *     and with regex.findall on fuzz strings (tests/test_oracle_vs_reference.py).
 *
 * The pre-tokeniser regex engine lives in the third-party `regex` module
 * (pinned regex==2025.11.3 in the reference's uv.lock:279-280, 2026.3.32 installed);
 * its Unicode classes are dumped into unicode_tables.inc by tools/gen_unicode_tables.py
 * and its matching semantics for the one pattern used
 *     sp1|sp2|...|'(?:[sdmt]|ll|ve|re)| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+
 * are restated below as a sequential scanner.
 *
 * All citations are relative to /root/reference/.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#include "unicode_tables.inc"

#define CLS_O 0
#define CLS_L 1
#define CLS_N 2
#define CLS_S 3

/* ------------------------------------------------------------------------- */
/* UTF-8 helpers                                                               */
/* ------------------------------------------------------------------------- */

static inline int cls_of(uint32_t cp) {
    unsigned blk = yabpe_ucd_stage1[cp >> 8];
    unsigned byte = yabpe_ucd_stage2[blk * 64 + ((cp & 255) >> 2)];
    return (byte >> ((cp & 3) * 2)) & 3;
}

int orc_class_of(uint32_t cp) { return cp < 0x110000 ? cls_of(cp) : 0; }

/* decode the (already validated) code point starting at p; *len = its byte length */
static inline uint32_t dec_cp(const uint8_t *p, int *len) {
    uint8_t b = p[0];
    if (b < 0x80) { *len = 1; return b; }
    if (b < 0xE0) { *len = 2; return ((uint32_t)(b & 0x1F) << 6) | (p[1] & 0x3F); }
    if (b < 0xF0) { *len = 3; return ((uint32_t)(b & 0x0F) << 12) | ((uint32_t)(p[1] & 0x3F) << 6) | (p[2] & 0x3F); }
    *len = 4;
    return ((uint32_t)(b & 0x07) << 18) | ((uint32_t)(p[1] & 0x3F) << 12) | ((uint32_t)(p[2] & 0x3F) << 6) | (p[3] & 0x3F);
}

/*
 * Strict UTF-8 validation with CPython's bytes.decode('utf-8') acceptance set
 * (trainer.py:156-160: the ValueError carries start + e.start).  Returns the byte
 * offset of the first ill-formed sequence's first byte, or -1 when valid.
 */
int64_t orc_utf8_first_error(const uint8_t *d, int64_t n) {
    int64_t i = 0;
    while (i < n) {
        uint8_t b = d[i];
        if (b < 0x80) { i++; continue; }
        int need; uint8_t lo = 0x80, hi = 0xBF;
        if (b >= 0xC2 && b <= 0xDF) need = 1;
        else if (b == 0xE0) { need = 2; lo = 0xA0; }
        else if ((b >= 0xE1 && b <= 0xEC) || b == 0xEE || b == 0xEF) need = 2;
        else if (b == 0xED) { need = 2; hi = 0x9F; }
        else if (b == 0xF0) { need = 3; lo = 0x90; }
        else if (b >= 0xF1 && b <= 0xF3) need = 3;
        else if (b == 0xF4) { need = 3; hi = 0x8F; }
        else return i;
        for (int k = 1; k <= need; k++) {
            if (i + k >= n) return i;
            uint8_t c = d[i + k];
            if (k == 1) { if (c < lo || c > hi) return i; }
            else if (c < 0x80 || c > 0xBF) return i;
        }
        i += need + 1;
    }
    return -1;
}

/* ------------------------------------------------------------------------- */
/* P1: reference chunk cuts  (trainer.py:139-144, 172-198)                     */
/* ------------------------------------------------------------------------- */

/* Writes chunk END offsets (exclusive) into cuts[]; returns the number of chunks. */
int64_t orc_chunk_cuts(const uint8_t *d, int64_t n, int64_t chunk_size, int64_t *cuts, int64_t cap) {
    if (n == 0) return 0;
    if (n <= chunk_size) { if (cap > 0) cuts[0] = n; return 1; }
    int64_t start = 0, k = 0;
    while (start < n) {
        int64_t tentative = start + chunk_size < n ? start + chunk_size : n;
        int64_t actual;
        if (tentative < n) {
            int64_t bstart = tentative - 4 > 0 ? tentative - 4 : 0;
            int64_t pos = tentative - bstart; /* index into d[bstart .. tentative] */
            int64_t blen = tentative + 1 - bstart;
            if (pos >= blen) pos = blen;
            else while (pos > 0 && (d[bstart + pos] & 0xC0) == 0x80) pos--;
            actual = bstart + pos;
        } else actual = n;
        if (actual > start) { if (k < cap) cuts[k] = actual; k++; start = actual; }
        else start += 1;
    }
    return k;
}

/* ------------------------------------------------------------------------- */
/* P2: sequential pre-tokeniser                                                */
/* ------------------------------------------------------------------------- */

typedef struct {
    const uint8_t *blob;     /* concatenated special-token bytes, priority order */
    const int32_t *offs;     /* n+1 offsets                                       */
    int32_t n;
} orc_specials;

static inline int special_at(const orc_specials *sp, const uint8_t *d, int64_t i, int64_t end) {
    for (int s = 0; s < sp->n; s++) {
        int32_t len = sp->offs[s + 1] - sp->offs[s];
        if (len > 0 && i + len <= end && memcmp(d + i, sp->blob + sp->offs[s], (size_t)len) == 0) return s;
    }
    return -1;
}

typedef void (*orc_emit_fn)(void *ctx, int64_t start, int64_t end, int special);

/* GPT-2 pattern only, on d[start,end) as one independent text. */
static void scan_gpt2(const uint8_t *d, int64_t start, int64_t end,
                      const orc_specials *sp_trainer, orc_emit_fn emit, void *ctx) {
    int64_t i = start;
    while (i < end) {
        /* (1) trainer mode: specials are the first alternatives, list order (trainer.py:165-167) */
        if (sp_trainer) {
            int s = special_at(sp_trainer, d, i, end);
            if (s >= 0) {
                int64_t e = i + (sp_trainer->offs[s + 1] - sp_trainer->offs[s]);
                emit(ctx, i, e, s);
                i = e;
                continue;
            }
        }
        /* (2) '(?:[sdmt]|ll|ve|re) */
        if (d[i] == '\'' && i + 1 < end) {
            uint8_t c1 = d[i + 1];
            int clen = 0;
            if (c1 == 's' || c1 == 'd' || c1 == 'm' || c1 == 't') clen = 2;
            else if (i + 2 < end) {
                uint8_t c2 = d[i + 2];
                if ((c1 == 'l' && c2 == 'l') || (c1 == 'v' && c2 == 'e') || (c1 == 'r' && c2 == 'e')) clen = 3;
            }
            if (clen) { emit(ctx, i, i + clen, -1); i += clen; continue; }
        }
        /* (3)  ?\p{L}+ |  ?\p{N}+ |  ?[^\s\p{L}\p{N}]+ */
        int64_t j = i; int len; uint32_t cp;
        if (d[i] == ' ' && i + 1 < end) {
            cp = dec_cp(d + i + 1, &len);
            if (cls_of(cp) != CLS_S) j = i + 1;
        }
        cp = dec_cp(d + j, &len);
        int c = cls_of(cp);
        if (c != CLS_S) {
            int64_t e = j + len;
            while (e < end) {
                cp = dec_cp(d + e, &len);
                if (cls_of(cp) != c) break;
                e += len;
            }
            emit(ctx, i, e, -1);
            i = e;
            continue;
        }
        /* (4) \s+(?!\S) | \s+ */
        int64_t e = i, last = i; int ncp = 0;
        while (e < end) {
            cp = dec_cp(d + e, &len);
            if (cls_of(cp) != CLS_S) break;
            last = e; e += len; ncp++;
        }
        if (e == end || ncp == 1) { emit(ctx, i, e, -1); i = e; }
        else { emit(ctx, i, last, -1); i = last; }
    }
}

/*
 * Encode-mode scan (tokenizer.py:97-102,169-186): split at specials first
 * (leftmost, priority order = longest first as sorted by the caller), then the
 * plain GPT-2 pattern on every part as an independent text.
 */
static void scan_encode(const uint8_t *d, int64_t start, int64_t end,
                        const orc_specials *sp, orc_emit_fn emit, void *ctx) {
    int64_t part = start, i = start;
    if (sp && sp->n > 0) {
        while (i < end) {
            int s = special_at(sp, d, i, end);
            if (s >= 0) {
                if (i > part) scan_gpt2(d, part, i, NULL, emit, ctx);
                int64_t e = i + (sp->offs[s + 1] - sp->offs[s]);
                emit(ctx, i, e, s);
                i = e; part = e;
            } else i++;
        }
    }
    if (end > part) scan_gpt2(d, part, end, NULL, emit, ctx);
}

typedef struct { int64_t *starts; int32_t *kinds; int64_t cap, n; } collect_ctx;
static void collect_emit(void *vctx, int64_t s, int64_t e, int special) {
    collect_ctx *c = (collect_ctx *)vctx; (void)e;
    if (c->n < c->cap) { c->starts[c->n] = s; if (c->kinds) c->kinds[c->n] = special; }
    c->n++;
}

/*
 * mode 0 = trainer (specials as leading alternatives), 1 = encode (split first).
 * `cuts` = chunk END offsets (P1); every chunk is an independent text.
 * Returns the number of tokens; starts[k] / kinds[k] filled up to cap.  Token k
 * ends at starts[k+1], or at its chunk end (cuts are always token starts).
 */
int64_t orc_pretokenize(const uint8_t *d, int64_t n, const int64_t *cuts, int64_t n_cuts,
                        const uint8_t *sp_blob, const int32_t *sp_offs, int32_t n_sp, int mode,
                        int64_t *starts, int32_t *kinds, int64_t cap) {
    orc_specials sp = { sp_blob, sp_offs, n_sp };
    collect_ctx c = { starts, kinds, cap, 0 };
    int64_t s = 0;
    for (int64_t k = 0; k <= n_cuts; k++) {
        int64_t e = k < n_cuts ? cuts[k] : n;
        if (e > s) {
            if (mode == 0) scan_gpt2(d, s, e, n_sp > 0 ? &sp : NULL, collect_emit, &c);
            else scan_encode(d, s, e, &sp, collect_emit, &c);
        }
        s = e;
    }
    return c.n;
}

/* ------------------------------------------------------------------------- */
/* small containers                                                            */
/* ------------------------------------------------------------------------- */

static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
static uint64_t hash_bytes(const uint8_t *p, int64_t n) {
    uint64_t h = 0xcbf29ce484222325ULL ^ (uint64_t)n;
    for (int64_t i = 0; i < n; i++) { h ^= p[i]; h *= 0x100000001b3ULL; }
    return mix64(h);
}

/* byte-string interner: bytes -> dense id (used for words and for tokens) */
typedef struct {
    uint8_t *pool; int64_t pool_len, pool_cap;
    int64_t *off; int32_t *len; int64_t n, cap;
    int64_t *slots; int64_t nslots;   /* -1 empty, else id */
} interner;

static void in_init(interner *t) {
    memset(t, 0, sizeof *t);
    t->pool_cap = 1 << 16; t->pool = (uint8_t *)malloc((size_t)t->pool_cap);
    t->cap = 1024; t->off = (int64_t *)malloc(sizeof(int64_t) * t->cap); t->len = (int32_t *)malloc(sizeof(int32_t) * t->cap);
    t->nslots = 4096; t->slots = (int64_t *)malloc(sizeof(int64_t) * t->nslots);
    for (int64_t i = 0; i < t->nslots; i++) t->slots[i] = -1;
}
static void in_free(interner *t) { free(t->pool); free(t->off); free(t->len); free(t->slots); }
static void in_rehash(interner *t) {
    int64_t ns = t->nslots * 2;
    int64_t *s = (int64_t *)malloc(sizeof(int64_t) * ns);
    for (int64_t i = 0; i < ns; i++) s[i] = -1;
    for (int64_t id = 0; id < t->n; id++) {
        uint64_t h = hash_bytes(t->pool + t->off[id], t->len[id]) & (uint64_t)(ns - 1);
        while (s[h] >= 0) h = (h + 1) & (uint64_t)(ns - 1);
        s[h] = id;
    }
    free(t->slots); t->slots = s; t->nslots = ns;
}
static int64_t in_find(const interner *t, const uint8_t *p, int64_t n) {
    uint64_t h = hash_bytes(p, n) & (uint64_t)(t->nslots - 1);
    while (t->slots[h] >= 0) {
        int64_t id = t->slots[h];
        if (t->len[id] == n && memcmp(t->pool + t->off[id], p, (size_t)n) == 0) return id;
        h = (h + 1) & (uint64_t)(t->nslots - 1);
    }
    return -1;
}
static int64_t in_add(interner *t, const uint8_t *p, int64_t n, int *is_new) {
    int64_t id = in_find(t, p, n);
    if (id >= 0) { if (is_new) *is_new = 0; return id; }
    if (is_new) *is_new = 1;
    if (t->pool_len + n > t->pool_cap) {
        while (t->pool_len + n > t->pool_cap) t->pool_cap *= 2;
        t->pool = (uint8_t *)realloc(t->pool, (size_t)t->pool_cap);
    }
    if (t->n == t->cap) {
        t->cap *= 2;
        t->off = (int64_t *)realloc(t->off, sizeof(int64_t) * t->cap);
        t->len = (int32_t *)realloc(t->len, sizeof(int32_t) * t->cap);
    }
    id = t->n++;
    memcpy(t->pool + t->pool_len, p, (size_t)n);
    t->off[id] = t->pool_len; t->len[id] = (int32_t)n; t->pool_len += n;
    if (t->n * 2 > t->nslots) in_rehash(t);
    else {
        uint64_t h = hash_bytes(p, n) & (uint64_t)(t->nslots - 1);
        while (t->slots[h] >= 0) h = (h + 1) & (uint64_t)(t->nslots - 1);
        t->slots[h] = id;
    }
    return id;
}

/* ------------------------------------------------------------------------- */
/* trainer                                                                     */
/* ------------------------------------------------------------------------- */

typedef struct { int32_t *v; int64_t n, cap; } ivec;
static void iv_push(ivec *a, int32_t x) {
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 4; a->v = (int32_t *)realloc(a->v, sizeof(int32_t) * a->cap); }
    a->v[a->n++] = x;
}

typedef struct {
    uint64_t key;      /* (a << 32) | b, token ids */
    int64_t count;     /* pair_counts[pair]; entry is "absent" when count <= 0 (trainer.py:268-271) */
    ivec words;        /* pair_to_words[pair] (superset; stale entries are re-checked) */
    int64_t heap_stamp;
} pair_ent;

typedef struct orc_trainer {
    orc_specials sp; uint8_t *sp_blob; int32_t *sp_offs;
    interner words;          /* unique pre-token byte strings (word_freq keys, trainer.py:221-225) */
    int64_t *word_freq; int64_t word_freq_cap;
    int64_t n_pretokens;
    /* merge-loop state */
    interner toks;           /* vocab: token bytes -> id (trainer.py:119-134) */
    int32_t **wsym; int32_t *wlen;  /* current token sequence of each word */
    pair_ent *pairs; int64_t npairs, pairs_cap;
    int64_t *pslots; int64_t npslots;
    int32_t *merges; int64_t n_merges;   /* (a, b) token-id pairs */
    int32_t *merge_new;                  /* resulting token id (existing id when bytes were already in vocab) */
} orc_trainer;

orc_trainer *orc_trainer_new(const uint8_t *sp_blob, const int32_t *sp_offs, int32_t n_sp) {
    orc_trainer *t = (orc_trainer *)calloc(1, sizeof *t);
    int32_t tot = n_sp > 0 ? sp_offs[n_sp] : 0;
    t->sp_blob = (uint8_t *)malloc((size_t)tot + 1); if (tot) memcpy(t->sp_blob, sp_blob, (size_t)tot);
    t->sp_offs = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n_sp + 1));
    if (n_sp > 0) memcpy(t->sp_offs, sp_offs, sizeof(int32_t) * ((size_t)n_sp + 1)); else t->sp_offs[0] = 0;
    t->sp.blob = t->sp_blob; t->sp.offs = t->sp_offs; t->sp.n = n_sp;
    in_init(&t->words);
    t->word_freq_cap = 1024; t->word_freq = (int64_t *)calloc((size_t)t->word_freq_cap, sizeof(int64_t));
    return t;
}

void orc_trainer_free(orc_trainer *t) {
    if (!t) return;
    if (t->wsym) { for (int64_t w = 0; w < t->words.n; w++) free(t->wsym[w]); free(t->wsym); free(t->wlen); }
    if (t->pairs) { for (int64_t p = 0; p < t->npairs; p++) free(t->pairs[p].words.v); free(t->pairs); free(t->pslots); }
    if (t->toks.pool) in_free(&t->toks);
    free(t->merges); free(t->merge_new);
    in_free(&t->words); free(t->word_freq); free(t->sp_blob); free(t->sp_offs); free(t);
}

typedef struct { orc_trainer *t; const uint8_t *d; } count_ctx;
static void count_emit(void *vctx, int64_t s, int64_t e, int special) {
    count_ctx *c = (count_ctx *)vctx; (void)special;   /* specials are ordinary words */
    orc_trainer *t = c->t;
    if (e <= s) return;
    int64_t id = in_add(&t->words, c->d + s, e - s, NULL);
    if (id >= t->word_freq_cap) {
        int64_t nc = t->word_freq_cap * 2;
        t->word_freq = (int64_t *)realloc(t->word_freq, sizeof(int64_t) * nc);
        memset(t->word_freq + t->word_freq_cap, 0, sizeof(int64_t) * (nc - t->word_freq_cap));
        t->word_freq_cap = nc;
    }
    t->word_freq[id]++;
    t->n_pretokens++;
}

/*
 * One input file: P1 chunking, strict UTF-8 per chunk, P2 scan, P5 word counts.
 * Returns 0, or -1 with *err_pos = byte offset of the first invalid UTF-8 sequence
 * (the reference raises ValueError there; chunks are decoded in order).
 */
int orc_trainer_feed(orc_trainer *t, const uint8_t *d, int64_t n, int64_t chunk_size, int64_t *err_pos) {
    if (n == 0) return 0;
    int64_t ncuts = orc_chunk_cuts(d, n, chunk_size, NULL, 0);
    int64_t *cuts = (int64_t *)malloc(sizeof(int64_t) * (size_t)(ncuts + 1));
    orc_chunk_cuts(d, n, chunk_size, cuts, ncuts);
    int64_t s = 0;
    for (int64_t k = 0; k < ncuts; k++) {
        int64_t e = cuts[k];
        int64_t bad = orc_utf8_first_error(d + s, e - s);
        if (bad >= 0) { if (err_pos) *err_pos = s + bad; free(cuts); return -1; }
        s = e;
    }
    count_ctx c = { t, d };
    s = 0;
    for (int64_t k = 0; k < ncuts; k++) {
        scan_gpt2(d, s, cuts[k], t->sp.n > 0 ? &t->sp : NULL, count_emit, &c);
        s = cuts[k];
    }
    free(cuts);
    return 0;
}

/* feed already-split pre-tokens (mirror of _merge_loop(sequences), trainer.py:216) */
void orc_trainer_feed_word(orc_trainer *t, const uint8_t *w, int64_t n, int64_t freq) {
    count_ctx c = { t, w };
    if (n <= 0) return;
    count_emit(&c, 0, n, -1);
    int64_t id = in_find(&t->words, w, n);
    t->word_freq[id] += freq - 1;
    t->n_pretokens += freq - 1;
}

int64_t orc_trainer_num_words(const orc_trainer *t) { return t->words.n; }
int64_t orc_trainer_num_pretokens(const orc_trainer *t) { return t->n_pretokens; }
int64_t orc_trainer_words_bytes(const orc_trainer *t) { return t->words.pool_len; }
/* unique words in first-seen order: blob, offsets (n+1), freqs */
void orc_trainer_get_words(const orc_trainer *t, uint8_t *blob, int64_t *offs, int64_t *freqs) {
    memcpy(blob, t->words.pool, (size_t)t->words.pool_len);
    for (int64_t w = 0; w < t->words.n; w++) { offs[w] = t->words.off[w]; freqs[w] = t->word_freq[w]; }
    offs[t->words.n] = t->words.pool_len;
}

/* ---- pair table -------------------------------------------------------- */

static int64_t pair_find(orc_trainer *t, uint64_t key) {
    uint64_t h = mix64(key) & (uint64_t)(t->npslots - 1);
    while (t->pslots[h] >= 0) {
        if (t->pairs[t->pslots[h]].key == key) return t->pslots[h];
        h = (h + 1) & (uint64_t)(t->npslots - 1);
    }
    return -1;
}
static int64_t pair_get(orc_trainer *t, uint64_t key) {
    int64_t p = pair_find(t, key);
    if (p >= 0) return p;
    if (t->npairs == t->pairs_cap) {
        t->pairs_cap *= 2;
        t->pairs = (pair_ent *)realloc(t->pairs, sizeof(pair_ent) * (size_t)t->pairs_cap);
    }
    p = t->npairs++;
    memset(&t->pairs[p], 0, sizeof(pair_ent));
    t->pairs[p].key = key;
    if (t->npairs * 2 > t->npslots) {
        int64_t ns = t->npslots * 2;
        free(t->pslots); t->pslots = (int64_t *)malloc(sizeof(int64_t) * (size_t)ns);
        for (int64_t i = 0; i < ns; i++) t->pslots[i] = -1;
        t->npslots = ns;
        for (int64_t q = 0; q < t->npairs; q++) {
            uint64_t h = mix64(t->pairs[q].key) & (uint64_t)(ns - 1);
            while (t->pslots[h] >= 0) h = (h + 1) & (uint64_t)(ns - 1);
            t->pslots[h] = q;
        }
    } else {
        uint64_t h = mix64(key) & (uint64_t)(t->npslots - 1);
        while (t->pslots[h] >= 0) h = (h + 1) & (uint64_t)(t->npslots - 1);
        t->pslots[h] = p;
    }
    return p;
}

/* Python bytes ordering: lexicographic unsigned, a proper prefix is smaller. */
static inline int tok_cmp(const interner *tk, int32_t x, int32_t y) {
    if (x == y) return 0;
    int32_t lx = tk->len[x], ly = tk->len[y];
    int r = memcmp(tk->pool + tk->off[x], tk->pool + tk->off[y], (size_t)(lx < ly ? lx : ly));
    if (r) return r;
    return lx < ly ? -1 : (lx > ly ? 1 : 0);
}
/* key = (count, (left_bytes, right_bytes))  -- trainer.py:246 */
static inline int pair_gt(const orc_trainer *t, const pair_ent *p, const pair_ent *q) {
    if (p->count != q->count) return p->count > q->count;
    int r = tok_cmp(&t->toks, (int32_t)(p->key >> 32), (int32_t)(q->key >> 32));
    if (r) return r > 0;
    return tok_cmp(&t->toks, (int32_t)(p->key & 0xffffffffu), (int32_t)(q->key & 0xffffffffu)) > 0;
}

/* lazy max-heap used only by fast mode */
typedef struct { int64_t pair; int64_t count; } hent;
typedef struct { hent *v; int64_t n, cap; } heap_t;
static int hent_gt(const orc_trainer *t, hent a, hent b) {
    if (a.count != b.count) return a.count > b.count;
    if (a.pair == b.pair) return 0;
    pair_ent pa = t->pairs[a.pair], pb = t->pairs[b.pair];
    pa.count = pb.count = 0;
    return pair_gt(t, &pa, &pb);
}
static void heap_push(const orc_trainer *t, heap_t *h, hent e) {
    if (h->n == h->cap) { h->cap = h->cap ? h->cap * 2 : 1024; h->v = (hent *)realloc(h->v, sizeof(hent) * (size_t)h->cap); }
    int64_t i = h->n++;
    while (i > 0) {
        int64_t p = (i - 1) / 2;
        if (!hent_gt(t, e, h->v[p])) break;
        h->v[i] = h->v[p]; i = p;
    }
    h->v[i] = e;
}
static hent heap_pop(const orc_trainer *t, heap_t *h) {
    hent top = h->v[0], e = h->v[--h->n];
    int64_t i = 0;
    for (;;) {
        int64_t l = 2 * i + 1, r = l + 1, m = l;
        if (l >= h->n) break;
        if (r < h->n && hent_gt(t, h->v[r], h->v[l])) m = r;
        if (!hent_gt(t, h->v[m], e)) break;
        h->v[i] = h->v[m]; i = m;
    }
    if (h->n > 0) h->v[i] = e;
    return top;
}

/*
 * _merge_loop (trainer.py:216-302).  `fast` != 0 replaces the O(|pairs|) max()
 * scan of trainer.py:246 by a lazy heap with the same ordering; the default (0)
 * keeps the reference's linear scan so that the port's cost model matches.
 * Returns the number of merges performed.
 */
int64_t orc_trainer_run(orc_trainer *t, int64_t vocab_size, int64_t min_frequency, int fast) {
    /* _init_base_vocab (trainer.py:119-134) */
    in_init(&t->toks);
    for (int b = 0; b < 256; b++) { uint8_t x = (uint8_t)b; in_add(&t->toks, &x, 1, NULL); }
    for (int s = 0; s < t->sp.n; s++)
        in_add(&t->toks, t->sp.blob + t->sp.offs[s], t->sp.offs[s + 1] - t->sp.offs[s], NULL);
    int64_t nw = t->words.n;
    if (nw == 0) return 0;               /* trainer.py:81-85 */

    t->wsym = (int32_t **)malloc(sizeof(int32_t *) * (size_t)nw);
    t->wlen = (int32_t *)malloc(sizeof(int32_t) * (size_t)nw);
    t->pairs_cap = 1024; t->pairs = (pair_ent *)malloc(sizeof(pair_ent) * (size_t)t->pairs_cap);
    t->npslots = 4096; t->pslots = (int64_t *)malloc(sizeof(int64_t) * (size_t)t->npslots);
    for (int64_t i = 0; i < t->npslots; i++) t->pslots[i] = -1;

    /* word_freq / pair_counts / pair_to_words (trainer.py:221-235) */
    for (int64_t w = 0; w < nw; w++) {
        int32_t n = t->words.len[w];
        const uint8_t *p = t->words.pool + t->words.off[w];
        t->wsym[w] = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
        t->wlen[w] = n;
        for (int32_t j = 0; j < n; j++) t->wsym[w][j] = p[j];
        for (int32_t j = 0; j + 1 < n; j++) {
            int64_t pe = pair_get(t, ((uint64_t)p[j] << 32) | p[j + 1]);
            t->pairs[pe].count += t->word_freq[w];
            ivec *v = &t->pairs[pe].words;
            if (v->n == 0 || v->v[v->n - 1] != (int32_t)w) iv_push(v, (int32_t)w);
        }
    }

    int64_t num_merges = vocab_size - t->toks.n; if (num_merges < 0) num_merges = 0;   /* trainer.py:238 */
    t->merges = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(num_merges + 1));
    t->merge_new = (int32_t *)malloc(sizeof(int32_t) * (size_t)(num_merges + 1));
    t->n_merges = 0;

    heap_t heap = { 0, 0, 0 };
    if (fast) for (int64_t p = 0; p < t->npairs; p++) { hent e = { p, t->pairs[p].count }; heap_push(t, &heap, e); }
    int64_t *stamp = (int64_t *)calloc((size_t)nw, sizeof(int64_t));
    uint8_t *tmp = NULL; int64_t tmp_cap = 0;
    ivec touched = { 0, 0, 0 };

    for (int64_t m = 0; m < num_merges; m++) {
        /* best pair: max over (count, (left bytes, right bytes))  trainer.py:242-248 */
        int64_t best = -1;
        if (!fast) {
            for (int64_t p = 0; p < t->npairs; p++) {
                if (t->pairs[p].count <= 0) continue;
                if (best < 0 || pair_gt(t, &t->pairs[p], &t->pairs[best])) best = p;
            }
        } else {
            while (heap.n > 0) {
                hent e = heap_pop(t, &heap);
                if (t->pairs[e.pair].count == e.count && e.count > 0) { best = e.pair; break; }
            }
        }
        if (best < 0) break;                                   /* `if not pair_counts: break` */
        if (t->pairs[best].count < min_frequency) break;
        int32_t a = (int32_t)(t->pairs[best].key >> 32), b = (int32_t)(t->pairs[best].key & 0xffffffffu);

        /* merged = p0 + p1; id only if bytes are new (trainer.py:250-251, 298-300) */
        int64_t la = t->toks.len[a], lb = t->toks.len[b];
        if (la + lb > tmp_cap) { tmp_cap = (la + lb) * 2; tmp = (uint8_t *)realloc(tmp, (size_t)tmp_cap); }
        memcpy(tmp, t->toks.pool + t->toks.off[a], (size_t)la);
        memcpy(tmp + la, t->toks.pool + t->toks.off[b], (size_t)lb);
        int32_t c = (int32_t)in_add(&t->toks, tmp, la + lb, NULL);

        /* affected words (trainer.py:254-294) */
        ivec aff = t->pairs[best].words;      /* take the list; new postings for `best` cannot appear */
        t->pairs[best].words.v = NULL; t->pairs[best].words.n = t->pairs[best].words.cap = 0;
        touched.n = 0;
        for (int64_t k = 0; k < aff.n; k++) {
            int32_t w = aff.v[k];
            if (stamp[w] == m + 1) continue;
            int32_t *s = t->wsym[w]; int32_t n = t->wlen[w];
            int has = 0;
            for (int32_t j = 0; j + 1 < n; j++) if (s[j] == a && s[j + 1] == b) { has = 1; break; }
            if (!has) continue;
            stamp[w] = m + 1;
            int64_t f = t->word_freq[w];
            for (int32_t j = 0; j + 1 < n; j++) {               /* old pairs: -freq */
                int64_t pe = pair_find(t, ((uint64_t)(uint32_t)s[j] << 32) | (uint32_t)s[j + 1]);
                t->pairs[pe].count -= f;
                if (fast) iv_push(&touched, (int32_t)pe);
            }
            int32_t o = 0;
            for (int32_t j = 0; j < n;) {                         /* L->R, non-overlapping */
                if (j + 1 < n && s[j] == a && s[j + 1] == b) { s[o++] = c; j += 2; }
                else s[o++] = s[j++];
            }
            t->wlen[w] = o;
            for (int32_t j = 0; j + 1 < o; j++) {                /* new pairs: +freq, index */
                int64_t pe = pair_get(t, ((uint64_t)(uint32_t)s[j] << 32) | (uint32_t)s[j + 1]);
                t->pairs[pe].count += f;
                ivec *v = &t->pairs[pe].words;
                if (s[j] == c || s[j + 1] == c) { if (v->n == 0 || v->v[v->n - 1] != w) iv_push(v, w); }
                if (fast) iv_push(&touched, (int32_t)pe);
            }
        }
        free(aff.v);
        if (fast) for (int64_t k = 0; k < touched.n; k++) {
            int64_t pe = touched.v[k];
            if (t->pairs[pe].heap_stamp == m + 1) continue;
            t->pairs[pe].heap_stamp = m + 1;
            if (t->pairs[pe].count > 0) { hent e = { pe, t->pairs[pe].count }; heap_push(t, &heap, e); }
        }
        t->merges[2 * t->n_merges] = a; t->merges[2 * t->n_merges + 1] = b;
        t->merge_new[t->n_merges] = c;
        t->n_merges++;
    }
    free(stamp); free(tmp); free(touched.v); free(heap.v);
    return t->n_merges;
}

int64_t orc_trainer_vocab_size(const orc_trainer *t) { return t->toks.n; }
int64_t orc_trainer_vocab_bytes(const orc_trainer *t) { return t->toks.pool_len; }
void orc_trainer_get_vocab(const orc_trainer *t, uint8_t *blob, int64_t *offs) {
    memcpy(blob, t->toks.pool, (size_t)t->toks.pool_len);
    for (int64_t i = 0; i < t->toks.n; i++) offs[i] = t->toks.off[i];
    offs[t->toks.n] = t->toks.pool_len;
}
void orc_trainer_get_merges(const orc_trainer *t, int32_t *pairs, int32_t *newids) {
    memcpy(pairs, t->merges, sizeof(int32_t) * 2 * (size_t)t->n_merges);
    memcpy(newids, t->merge_new, sizeof(int32_t) * (size_t)t->n_merges);
}

/* ------------------------------------------------------------------------- */
/* tokenizer (encode)                                                          */
/* ------------------------------------------------------------------------- */

typedef struct orc_tok {
    orc_specials sp; uint8_t *sp_blob; int32_t *sp_offs; int32_t *sp_ids;   /* vocab id or -1 (dropped, tokenizer.py:179-181) */
    /* symbol = distinct byte string among single bytes + merge operands/results */
    int32_t byte_sym[256];
    int32_t *sym_out;        /* symbol -> vocab id (unk already substituted, tokenizer.py:297-306) */
    uint64_t *mkey; int32_t *mrank; int32_t *mres; int64_t nm_slots;      /* (a,b) -> rank, result symbol */
    int32_t single_unk;
} orc_tok;

orc_tok *orc_tok_new(int32_t n_syms, const int32_t *byte_sym, const int32_t *sym_out,
                     const int32_t *m_a, const int32_t *m_b, const int32_t *m_rank, const int32_t *m_res, int64_t n_m,
                     const uint8_t *sp_blob, const int32_t *sp_offs, const int32_t *sp_ids, int32_t n_sp) {
    orc_tok *t = (orc_tok *)calloc(1, sizeof *t);
    memcpy(t->byte_sym, byte_sym, sizeof t->byte_sym);
    t->sym_out = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_syms + 1));
    memcpy(t->sym_out, sym_out, sizeof(int32_t) * (size_t)n_syms);
    int64_t ns = 16; while (ns < n_m * 2 + 2) ns *= 2;
    t->nm_slots = ns;
    t->mkey = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)ns);
    t->mrank = (int32_t *)malloc(sizeof(int32_t) * (size_t)ns);
    t->mres = (int32_t *)malloc(sizeof(int32_t) * (size_t)ns);
    for (int64_t i = 0; i < ns; i++) t->mkey[i] = ~0ULL;
    for (int64_t i = 0; i < n_m; i++) {
        uint64_t key = ((uint64_t)(uint32_t)m_a[i] << 32) | (uint32_t)m_b[i];
        uint64_t h = mix64(key) & (uint64_t)(ns - 1);
        while (t->mkey[h] != ~0ULL && t->mkey[h] != key) h = (h + 1) & (uint64_t)(ns - 1);
        t->mkey[h] = key; t->mrank[h] = m_rank[i]; t->mres[h] = m_res[i];
    }
    int32_t tot = n_sp > 0 ? sp_offs[n_sp] : 0;
    t->sp_blob = (uint8_t *)malloc((size_t)tot + 1); if (tot) memcpy(t->sp_blob, sp_blob, (size_t)tot);
    t->sp_offs = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n_sp + 1));
    t->sp_ids = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n_sp + 1));
    if (n_sp > 0) { memcpy(t->sp_offs, sp_offs, sizeof(int32_t) * ((size_t)n_sp + 1)); memcpy(t->sp_ids, sp_ids, sizeof(int32_t) * (size_t)n_sp); }
    else t->sp_offs[0] = 0;
    t->sp.blob = t->sp_blob; t->sp.offs = t->sp_offs; t->sp.n = n_sp;
    return t;
}
void orc_tok_free(orc_tok *t) {
    if (!t) return;
    free(t->sym_out); free(t->mkey); free(t->mrank); free(t->mres); free(t->sp_blob); free(t->sp_offs); free(t->sp_ids); free(t);
}

typedef struct { orc_tok *t; const uint8_t *d; int32_t *out; int64_t cap, n; int32_t *buf; int64_t buf_cap; } enc_ctx;

static inline int merge_lookup(const orc_tok *t, int32_t a, int32_t b, int32_t *rank, int32_t *res) {
    uint64_t key = ((uint64_t)(uint32_t)a << 32) | (uint32_t)b;
    uint64_t h = mix64(key) & (uint64_t)(t->nm_slots - 1);
    while (t->mkey[h] != ~0ULL) {
        if (t->mkey[h] == key) { *rank = t->mrank[h]; *res = t->mres[h]; return 1; }
        h = (h + 1) & (uint64_t)(t->nm_slots - 1);
    }
    return 0;
}

/* _encode_word_impl (tokenizer.py:195-308): repeatedly merge the adjacent pair with
 * the smallest (rank, position); this is exactly what the heap + stale check pops. */
static void enc_emit(void *vctx, int64_t s, int64_t e, int special) {
    enc_ctx *c = (enc_ctx *)vctx; orc_tok *t = c->t;
    if (special >= 0) {
        if (t->sp_ids[special] >= 0) { if (c->n < c->cap) c->out[c->n] = t->sp_ids[special]; c->n++; }
        return;
    }
    int64_t n = e - s;
    if (n > c->buf_cap) { c->buf_cap = n * 2; c->buf = (int32_t *)realloc(c->buf, sizeof(int32_t) * (size_t)c->buf_cap); }
    int32_t *w = c->buf;
    for (int64_t i = 0; i < n; i++) w[i] = t->byte_sym[c->d[s + i]];
    while (n > 1) {
        int32_t best_rank = INT32_MAX, best_res = 0; int64_t best_pos = -1;
        for (int64_t i = 0; i + 1 < n; i++) {
            int32_t r, res;
            if (merge_lookup(t, w[i], w[i + 1], &r, &res) && r < best_rank) { best_rank = r; best_res = res; best_pos = i; }
        }
        if (best_pos < 0) break;
        w[best_pos] = best_res;
        memmove(w + best_pos + 1, w + best_pos + 2, sizeof(int32_t) * (size_t)(n - best_pos - 2));
        n--;
    }
    for (int64_t i = 0; i < n; i++) { if (c->n < c->cap) c->out[c->n] = t->sym_out[w[i]]; c->n++; }
}

/* encode one text (tokenizer.py:152-193); returns the id count (fills up to cap) */
int64_t orc_tok_encode(orc_tok *t, const uint8_t *d, int64_t n, int32_t *out, int64_t cap) {
    enc_ctx c = { t, d, out, cap, 0, NULL, 0 };
    if (n > 0) scan_encode(d, 0, n, &t->sp, enc_emit, &c);
    free(c.buf);
    return c.n;
}
