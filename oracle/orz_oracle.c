/*
 * orz_oracle.c — CPU restatement of richox/orz's ROLZ stream codec.
 *
 * TEST INFRASTRUCTURE ONLY (see orz_oracle.h).  Plain C11, single thread per stream exactly like
 * the reference.  All file:line citations are into the richox/orz tree @29414c6.
 *
 * Parity status: PARITY UNPINNED against real `orz` output.  The reference holds no golden byte vectors (its tests are
 * round-trips, coder.rs:219-266 and CI) and cannot be built here (nightly Rust, un-vendored crates).  What pins this
 * restatement: coder.rs:219-266 re-expressed, the hand-derived vectors of SURVEY App. D (reproduced), byte-exact round
 * trips at every level, and the whole-enwik8 sizes — 27,863,045 / 27,182,097 / 26,856,813 bytes at -l0/-l1/-l2, which
 * are 0.13 % BELOW the README's 27,898,433 / 27,217,825 / 26,892,825 (README.md:43-46; the README names tag v1.6.2,
 * this commit is crate 1.6.1).  See DESIGN.md §1.
 */
#include "orz_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * mem.rs:9-37 — unaligned little-endian get/put
 * ---------------------------------------------------------------------------------------- */
static inline uint32_t get_u32(const uint8_t *buf, size_t pos) {
    uint32_t v;
    memcpy(&v, buf + pos, 4);
    return v;
}
static inline void put_u32(uint8_t *buf, size_t pos, uint32_t v) { memcpy(buf + pos, &v, 4); }

/* mem.rs:41-51 — LCP in 16-byte steps, capped at max_len (a multiple of 16).  Two 8-byte compares per step stand in
 * for the reference's u8x16 compare (same reads, same result), so that the CPU baseline is not handicapped. */
static inline size_t mem_fast_common_prefix(const uint8_t *buf, size_t p1, size_t p2, size_t max_len) {
    for (size_t l = 0; l < max_len; l += 16) {
        uint64_t a0, a1, b0, b1;
        memcpy(&a0, buf + p1 + l, 8);
        memcpy(&a1, buf + p1 + l + 8, 8);
        memcpy(&b0, buf + p2 + l, 8);
        memcpy(&b1, buf + p2 + l + 8, 8);
        const uint64_t x0 = a0 ^ b0, x1 = a1 ^ b1;
        if (x0) return l + ((size_t)__builtin_ctzll(x0) >> 3);
        if (x1) return l + 8 + ((size_t)__builtin_ctzll(x1) >> 3);
    }
    return max_len;
}

/* mem.rs:55-70 — compares [len-4,len) then dwords at len-4-.. step 4 downwards; covers exactly [0,len) for len>=4 */
static inline int mem_fast_equal(const uint8_t *buf, size_t p1, size_t p2, size_t len, uint32_t p2_last_dword) {
    if (p2_last_dword != get_u32(buf, p1 + len - 4)) return 0;
    /* (0..len-4).step_by(4).rev() */
    if (len > 4) {
        size_t last = ((len - 4 - 1) / 4) * 4;
        for (size_t l = last;; l -= 4) {
            if (get_u32(buf, p1 + l) != get_u32(buf, p2 + l)) return 0;
            if (l == 0) break;
        }
    }
    return 1;
}

/* mem.rs:74-92 — LZ overlapped copy in 4-byte granules (scribbles a few bytes past the end) */
static inline void mem_fast_copy(uint8_t *buf, size_t psrc, size_t pdst, size_t len) {
    if (len == 4 && psrc + 4 <= pdst) {
        put_u32(buf, pdst, get_u32(buf, psrc));
        return;
    }
    while (pdst - psrc < 4) {
        put_u32(buf, pdst, get_u32(buf, psrc));
        pdst += pdst - psrc;
    }
    for (size_t l = 0; l < len; l += 4) put_u32(buf, pdst + l, get_u32(buf, psrc + l));
}

/* ------------------------------------------------------------------------------------------
 * instrumentation
 * ---------------------------------------------------------------------------------------- */
static orz_oracle_stats g_stats;
static int g_stats_on = 0;
void orz_oracle_stats_reset(void) {
    memset(&g_stats, 0, sizeof g_stats);
    g_stats_on = 1;
}
void orz_oracle_stats_get(orz_oracle_stats *s) { *s = g_stats; }
#define STAT(x) do { if (g_stats_on) { x; } } while (0)

/* ------------------------------------------------------------------------------------------
 * matcher.rs — Bucket (offset ring) and BucketMatcher (hash chains)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t nodes1[ORZ_BUCKET_ITEMS]; /* pos:25 (low) | match_len_min:7  (matcher.rs:231-237) */
    uint8_t nodes2[ORZ_BUCKET_ITEMS];  /* match_len_expected               (matcher.rs:239-243) */
    uint32_t head;
} bucket_t;

typedef struct {
    int16_t heads[ORZ_BUCKET_HASH];
    int16_t nexts[ORZ_BUCKET_ITEMS];
} matcher_t;

#define NODE_POS(n) ((n) & 0x1ffffffu)
#define NODE_LEN_MIN(n) ((n) >> 25)

static inline uint32_t node_add(uint32_t a, uint32_t b) { return (a + b) % ORZ_BUCKET_ITEMS; }                    /* matcher.rs:245-248 */
static inline uint32_t node_sub(uint32_t a, uint32_t b) { return (a + ORZ_BUCKET_ITEMS - b) % ORZ_BUCKET_ITEMS; } /* matcher.rs:250-253 */

/* matcher.rs:255-263 — sum_i (b[i]*MULS[i] ^ ADDS[i]) mod 2^32 */
uint32_t orz_oracle_hash_dword(const uint8_t *p) {
    return ((uint32_t)p[0] * 131313131u ^ 797u) + ((uint32_t)p[1] * 1313131u ^ 79797u) +
           ((uint32_t)p[2] * 13131u ^ 7979797u) + ((uint32_t)p[3] * 131u ^ 797979797u);
}

/* matcher.rs:54-60 */
static void bucket_init(bucket_t *b) { memset(b, 0, sizeof *b); }
/* matcher.rs:108-113 */
static void matcher_init(matcher_t *m) { memset(m, 0xff, sizeof *m); }

/* matcher.rs:62-80 */
static inline void bucket_update(bucket_t *b, size_t pos, size_t reduced_offset, size_t match_len) {
    uint32_t new_head = node_add(b->head, 1);
    if (match_len >= ORZ_LZ_MATCH_MIN_LEN) {
        uint32_t ni = node_sub(b->head, (uint32_t)reduced_offset);
        if (NODE_LEN_MIN(b->nodes1[ni]) <= match_len) {
            uint32_t lm = (uint32_t)(match_len + 1 < 127 ? match_len + 1 : 127);
            b->nodes1[ni] = NODE_POS(b->nodes1[ni]) | lm << 25;
        }
    }
    b->nodes1[new_head] = (uint32_t)pos & 0x1ffffffu;
    b->nodes2[new_head] = (uint8_t)match_len;
    b->head = new_head;
}

/* matcher.rs:82-87 */
static void bucket_forward(bucket_t *b, size_t forward_len) {
    for (int i = 0; i < ORZ_BUCKET_ITEMS; i++) {
        uint32_t n = b->nodes1[i];
        uint32_t p = NODE_POS(n);
        p = p > forward_len ? p - (uint32_t)forward_len : 0;
        b->nodes1[i] = p | (n & 0xfe000000u);
    }
}

/* matcher.rs:115-121 */
static inline void matcher_update(matcher_t *m, const bucket_t *b, const uint8_t *buf, size_t pos) {
    uint32_t head = b->head;
    uint32_t entry = orz_oracle_hash_dword(buf + pos) % ORZ_BUCKET_HASH;
    m->nexts[head] = m->heads[entry];
    m->heads[entry] = (int16_t)head;
}

/* matcher.rs:123-133 */
static void matcher_forward(matcher_t *m, const bucket_t *b) {
    for (int i = 0; i < ORZ_BUCKET_HASH; i++)
        if (m->heads[i] != -1 && NODE_POS(b->nodes1[m->heads[i]]) == 0) m->heads[i] = -1;
    for (int i = 0; i < ORZ_BUCKET_ITEMS; i++)
        if (m->nexts[i] != -1 && NODE_POS(b->nodes1[m->nexts[i]]) == 0) m->nexts[i] = -1;
}

typedef struct {
    size_t reduced_offset, match_len, match_len_expected, match_len_min;
} match_t;

/* matcher.rs:135-192 */
static inline match_t find_match(const matcher_t *m, const bucket_t *b, const uint8_t *buf, size_t buf_len, size_t pos,
                                 size_t match_depth) {
    match_t none = {0, 0, 0, 0};
    uint32_t entry = orz_oracle_hash_dword(buf + pos) % ORZ_BUCKET_HASH;
    int node_index = m->heads[entry];
    STAT(g_stats.fm_calls++);
    if (node_index == -1) return none;

    size_t max_len = ORZ_LZ_MATCH_MIN_LEN - 1;
    size_t max_match_len_min = ORZ_LZ_MATCH_MIN_LEN;
    size_t max_match_len_expected = ORZ_LZ_MATCH_MIN_LEN;
    int max_node_index = 0;
    size_t node_pos = NODE_POS(b->nodes1[node_index]);
    uint32_t max_len_dword = get_u32(buf, pos + max_len - 3);

    for (size_t it = 0; it < match_depth; it++) {
        uint32_t node_max_len_dword = get_u32(buf, node_pos + max_len - 3);
        STAT(g_stats.fm_hops++);
        if (node_max_len_dword == max_len_dword) {
            size_t lcp = mem_fast_common_prefix(buf, node_pos, pos, ORZ_LZ_MATCH_MAX_LEN);
            STAT(g_stats.fm_probe_hits++);
            if (lcp > max_len) {
                max_match_len_min = NODE_LEN_MIN(b->nodes1[node_index]);
                max_match_len_expected = b->nodes2[node_index];
                max_len = lcp;
                max_node_index = node_index;
                max_len_dword = get_u32(buf, pos + max_len - 3);
            }
            if (lcp == ORZ_LZ_MATCH_MAX_LEN) break;
            if (max_match_len_expected > 0 && lcp > max_match_len_expected) break;
        }
        node_index = m->nexts[node_index];
        if (node_index == -1) break;
        size_t node_pos_next = NODE_POS(b->nodes1[node_index]);
        if (node_pos <= node_pos_next) break;
        node_pos = node_pos_next;
    }

    if (max_len >= ORZ_LZ_MATCH_MIN_LEN && pos + max_len < buf_len) {
        match_t r;
        r.reduced_offset = node_sub(b->head, (uint32_t)max_node_index);
        r.match_len = max_len;
        r.match_len_expected = max_match_len_expected > ORZ_LZ_MATCH_MIN_LEN ? max_match_len_expected : ORZ_LZ_MATCH_MIN_LEN;
        r.match_len_min = max_match_len_min > ORZ_LZ_MATCH_MIN_LEN ? max_match_len_min : ORZ_LZ_MATCH_MIN_LEN;
        return r;
    }
    return none;
}

/* matcher.rs:194-228 — no end-of-buffer guard: reads bytes past buf_len (bit-visible) */
static inline int has_lazy_match(const matcher_t *m, const bucket_t *b, const uint8_t *buf, size_t pos,
                                 size_t min_match_len, size_t depth, uint64_t *hops) {
    uint32_t max_len_dword = get_u32(buf, pos + min_match_len - 4);
    uint32_t entry = orz_oracle_hash_dword(buf + pos) % ORZ_BUCKET_HASH;
    int node_index = m->heads[entry];
    if (node_index == -1) return 0;
    size_t node_pos = NODE_POS(b->nodes1[node_index]);
    for (size_t it = 0; it < depth; it++) {
        STAT((*hops)++);
        if (mem_fast_equal(buf, node_pos, pos, min_match_len, max_len_dword)) return 1;
        node_index = m->nexts[node_index];
        if (node_index == -1) break;
        size_t node_pos_next = NODE_POS(b->nodes1[node_index]);
        if (node_pos <= node_pos_next) break;
        node_pos = node_pos_next;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * symrank.rs — per-context symbol ranking coder
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint16_t value_array[ORZ_NUM_SYMBOLS];
    uint16_t index_array[ORZ_NUM_SYMBOLS];
    uint32_t encoded_cnt;
    uint32_t encoded_idx_sum;
} symrank_t;

/* symrank.rs:22-29 */
static void symrank_new(symrank_t *s) {
    memset(s, 0, sizeof *s);
    s->encoded_idx_sum = 1000000;
}
/* symrank.rs:31-36 */
static void symrank_init(symrank_t *s, const uint16_t *values) {
    for (int i = 0; i < ORZ_NUM_SYMBOLS; i++) {
        s->value_array[i] = values[i];
        s->index_array[values[i]] = (uint16_t)i;
    }
}
/* symrank.rs:61-97 */
static inline void symrank_update(symrank_t *s, uint16_t v, uint16_t i) {
    if (s->encoded_cnt > ORZ_NUM_SYMBOLS) {
        s->encoded_cnt = s->encoded_cnt * 9 / 10;
        s->encoded_idx_sum = s->encoded_idx_sum * 9 / 10;
    }
    s->encoded_cnt += 1;
    s->encoded_idx_sum += i;

    uint16_t sub = (uint16_t)(i / 16 + (uint16_t)(s->encoded_idx_sum / 16 / s->encoded_cnt));
    uint16_t next_i = i > sub ? (uint16_t)(i - sub) : 0; /* saturating_sub */
    if (next_i < i / 2) next_i = i / 2;                  /* .max(i / 2) */

    uint16_t d = (uint16_t)(i - next_i);
    if (d == 0) {
    } else if (d == 1) {
        uint16_t ni1 = next_i;
        uint16_t nv1 = s->value_array[ni1];
        s->index_array[v] = ni1;
        s->value_array[i] = nv1;
        s->index_array[nv1] = i;
        s->value_array[ni1] = v;
    } else {
        uint16_t ni2 = next_i;
        uint16_t ni1 = (uint16_t)(next_i + d / 2);
        uint16_t nv1 = s->value_array[ni1];
        uint16_t nv2 = s->value_array[ni2];
        s->value_array[i] = nv1;
        s->index_array[nv1] = i;
        s->value_array[ni1] = nv2;
        s->index_array[nv2] = ni1;
        s->value_array[ni2] = v;
        s->index_array[v] = ni2;
    }
}
/* symrank.rs:38-47 */
static inline uint16_t symrank_encode(symrank_t *s, uint16_t v, uint16_t vunlikely) {
    uint16_t i = s->index_array[v];
    uint16_t iun = s->index_array[vunlikely];
    symrank_update(s, v, i);
    if (i == iun) return ORZ_NUM_SYMBOLS - 1;
    return (uint16_t)(i - (i > iun));
}
/* symrank.rs:49-59 */
static inline uint16_t symrank_decode(symrank_t *s, uint16_t i, uint16_t vunlikely) {
    uint16_t iun = s->index_array[vunlikely];
    if (i == ORZ_NUM_SYMBOLS - 1) i = iun;
    else i = (uint16_t)(i + !(i < iun));
    uint16_t v = s->value_array[i];
    symrank_update(s, v, i);
    return v;
}

void orz_oracle_symrank_run(const uint16_t *init_order, const uint16_t *syms, const uint16_t *unlikely, size_t n,
                            uint16_t *ranks_out) {
    symrank_t s;
    symrank_new(&s);
    symrank_init(&s, init_order);
    for (size_t k = 0; k < n; k++) ranks_out[k] = symrank_encode(&s, syms[k], unlikely[k]);
}

/* ------------------------------------------------------------------------------------------
 * huffman.rs — static, length-limited canonical Huffman
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t weight;
    uint16_t index;
} heap_node;

/* min-heap on (weight, index): huffman.rs:28-38 reverses PartialOrd, which std's BinaryHeap uses.
 * Keys are unique, so pop order is fully determined whatever the heap's internal shape. */
static inline int heap_less(heap_node a, heap_node b) {
    return a.weight < b.weight || (a.weight == b.weight && a.index < b.index);
}
static void heap_push(heap_node *h, int *n, heap_node x) {
    int i = (*n)++;
    while (i > 0) {
        int p = (i - 1) / 2;
        if (!heap_less(x, h[p])) break;
        h[i] = h[p];
        i = p;
    }
    h[i] = x;
}
static heap_node heap_pop(heap_node *h, int *n) {
    heap_node top = h[0];
    heap_node x = h[--(*n)];
    int i = 0;
    for (;;) {
        int c = 2 * i + 1;
        if (c >= *n) break;
        if (c + 1 < *n && heap_less(h[c + 1], h[c])) c++;
        if (!heap_less(h[c], x)) break;
        h[i] = h[c];
        i = c;
    }
    if (*n > 0) h[i] = x;
    return top;
}

/* huffman.rs:27-111.  n <= 389. Returns max code length (0 for an empty table). */
int orz_oracle_huffman_lens(const uint32_t *sym_weights, int n, int max_code_len, uint8_t *lens) {
    enum { MAXN = 1024 };
    uint32_t weight[2 * MAXN];
    uint16_t child1[2 * MAXN], child2[2 * MAXN];
    uint8_t code_lens[2 * MAXN];
    heap_node heap[MAXN];
    for (int i = 0; i < n; i++) weight[i] = sym_weights[i];

    for (;;) {
        int hn = 0, nn = n;
        for (int i = 0; i < n; i++)
            if (sym_weights[i] > 0) {
                heap_node x = {weight[i], (uint16_t)i};
                heap_push(heap, &hn, x);
            }
        if (hn <= 1) { /* huffman.rs:64-71 */
            memset(lens, 0, (size_t)n);
            if (hn == 1) {
                lens[heap[0].index] = 1;
                return 1;
            }
            return 0;
        }
        while (hn > 1) { /* huffman.rs:73-88 */
            heap_node n1 = heap_pop(heap, &hn);
            heap_node n2 = heap_pop(heap, &hn);
            weight[nn] = n1.weight + n2.weight;
            child1[nn] = n1.index;
            child2[nn] = n2.index;
            heap_node x = {weight[nn], (uint16_t)nn};
            nn++;
            heap_push(heap, &hn, x);
        }
        memset(code_lens, 0, (size_t)nn); /* huffman.rs:91-96 */
        for (int i = nn - 1; i >= n; i--) {
            code_lens[child1[i]] = (uint8_t)(code_lens[i] + 1);
            code_lens[child2[i]] = (uint8_t)(code_lens[i] + 1);
        }
        int cur_max = 0;
        for (int i = 0; i < n; i++)
            if (code_lens[i] > cur_max) cur_max = code_lens[i];
        if (cur_max > max_code_len) { /* huffman.rs:99-108 */
            uint32_t shrink = 1u << (cur_max - max_code_len);
            for (int i = 0; i < n; i++)
                if (weight[i] > 0) {
                    uint32_t w = weight[i] / shrink;
                    weight[i] = w > 1 ? w : 1;
                }
            continue;
        }
        memcpy(lens, code_lens, (size_t)n);
        return cur_max;
    }
}

/* huffman.rs:119-141 — canonical codes in (len, sym) order */
void orz_oracle_huffman_codes(const uint8_t *lens, int n, uint16_t *codes) {
    uint16_t bits = 0;
    uint8_t cur = 1;
    memset(codes, 0, sizeof(uint16_t) * (size_t)n);
    for (int len = 1; len <= 16; len++) {
        for (int sym = 0; sym < n; sym++) {
            if (lens[sym] != len) continue;
            uint8_t shift = (uint8_t)(lens[sym] - cur);
            if ((int8_t)shift > 0) {
                bits = (uint16_t)(bits << shift);
                cur = (uint8_t)(cur + shift);
            }
            codes[sym] = bits;
            bits++;
        }
    }
}

/* huffman.rs:150-167 — flat LUT of (sym,len) indexed by the next max_code_len bits */
typedef struct {
    uint16_t sym[1 << 16];
    uint8_t len[1 << 16];
    uint8_t max_code_len;
} huff_dec_t;

static void huff_dec_build(huff_dec_t *d, const uint8_t *lens, int n, int max_code_len) {
    uint16_t codes[1024];
    orz_oracle_huffman_codes(lens, n, codes);
    d->max_code_len = (uint8_t)max_code_len;
    size_t size = (size_t)1 << max_code_len;
    memset(d->sym, 0, size * sizeof(uint16_t));
    memset(d->len, 0, size);
    for (int s = 0; s < n; s++) {
        if (lens[s] > 0) {
            int rest = max_code_len - lens[s];
            size_t base = (size_t)codes[s] << rest;
            for (size_t k = 0; k < ((size_t)1 << rest); k++) {
                d->sym[base + k] = (uint16_t)s;
                d->len[base + k] = lens[s];
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * coder.rs — MSB-first bit writer / reader over big-endian 32-bit words
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint64_t value;
    uint8_t len;
} bitbuf_t;

static inline uint64_t bb_peek(const bitbuf_t *b, uint8_t len) { /* coder.rs:166-169 */
    return (b->value >> (b->len - len)) & (((uint64_t)1 << len) - 1);
}
static inline void bb_put(bitbuf_t *b, uint8_t len, uint64_t value) { /* coder.rs:183-187 */
    b->value = (len >= 64 ? 0 : b->value << len) ^ value;
    b->len = (uint8_t)(b->len + len);
}
static inline uint64_t bb_get(bitbuf_t *b, uint8_t len) { /* coder.rs:176-181 */
    uint64_t v = bb_peek(b, len);
    b->len = (uint8_t)(b->len - len);
    return v;
}

typedef struct {
    uint8_t *out;
    size_t pos;
    bitbuf_t bb;
} bitenc_t;

static inline void enc_reserve(bitenc_t *e) { /* coder.rs:84-88, 198-206 */
    if (e->bb.len >= 32) {
        uint32_t w = (uint32_t)bb_get(&e->bb, 32);
        e->out[e->pos + 0] = (uint8_t)(w >> 24);
        e->out[e->pos + 1] = (uint8_t)(w >> 16);
        e->out[e->pos + 2] = (uint8_t)(w >> 8);
        e->out[e->pos + 3] = (uint8_t)w;
        e->pos += 4;
    }
}
static void enc_varint(bitenc_t *e, uint32_t v) { /* coder.rs:27-38 */
    for (;;) {
        enc_reserve(e);
        uint32_t has_next = v > 1;
        uint32_t bits = (v & 1) | has_next << 1;
        bb_put(&e->bb, 2, bits);
        v >>= 1;
        if (!has_next) break;
    }
}
static inline void enc_raw_bits(bitenc_t *e, uint32_t bits, uint8_t bits_len) { /* coder.rs:40-43 */
    enc_reserve(e);
    bb_put(&e->bb, bits_len, bits);
}
static void enc_huffman_table(bitenc_t *e, const uint8_t *lens, int n) { /* coder.rs:45-67 */
    int max_code_len = 0;
    for (int i = 0; i < n; i++)
        if (lens[i] > max_code_len) max_code_len = lens[i];
    enc_varint(e, (uint32_t)max_code_len);
    int last_sym = -1;
    for (int sym = 0; sym < n; sym++) {
        if (lens[sym] > 0) {
            uint32_t delta = last_sym < 0 ? (uint32_t)sym + 1 : (uint32_t)(sym - last_sym);
            enc_varint(e, delta);
            enc_varint(e, (uint32_t)(max_code_len - lens[sym]));
            last_sym = sym;
        }
    }
    enc_varint(e, 0);
}
static size_t enc_finish(bitenc_t *e) { /* coder.rs:75-82, 208-216 */
    enc_reserve(e);
    if (e->bb.len > 0) {
        bb_put(&e->bb, (uint8_t)(32 - e->bb.len), 0);
        while (e->bb.len > 0) {
            e->out[e->pos++] = (uint8_t)bb_peek(&e->bb, 8);
            e->bb.len = (uint8_t)(e->bb.len - (e->bb.len < 8 ? e->bb.len : 8));
        }
    }
    return e->pos;
}

typedef struct {
    const uint8_t *in;
    size_t pos;
    bitbuf_t bb;
} bitdec_t;

static inline void dec_reserve(bitdec_t *d) { /* coder.rs:152-156, 189-196: over-reads <=4 B past the payload */
    if (d->bb.len < 32) {
        uint32_t w = (uint32_t)d->in[d->pos] << 24 | (uint32_t)d->in[d->pos + 1] << 16 | (uint32_t)d->in[d->pos + 2] << 8 |
                     (uint32_t)d->in[d->pos + 3];
        bb_put(&d->bb, 32, w);
        d->pos += 4;
    }
}
static uint32_t dec_varint(bitdec_t *d) { /* coder.rs:108-120 */
    uint32_t v = 0;
    for (uint32_t bits_len = 0;; bits_len++) {
        dec_reserve(d);
        uint32_t bits = (uint32_t)bb_get(&d->bb, 2);
        if (bits_len < 32) v |= (bits & 1) << bits_len;
        if (!(bits > 1)) break;
    }
    return v;
}
static inline uint32_t dec_raw_bits(bitdec_t *d, uint8_t n) { /* coder.rs:122-125 */
    dec_reserve(d);
    return (uint32_t)bb_get(&d->bb, n);
}
/* coder.rs:127-142 — returns table length (last coded symbol + 1) */
static int dec_huffman_table(bitdec_t *d, uint8_t *lens, int cap, int *max_code_len) {
    *max_code_len = (uint8_t)dec_varint(d);
    int n = 0;
    for (;;) {
        uint32_t delta = dec_varint(d);
        if (delta == 0) break;
        for (uint32_t k = 1; k < delta; k++) {
            if (n >= cap) return -1;
            lens[n++] = 0;
        }
        if (n >= cap) return -1;
        lens[n++] = (uint8_t)(*max_code_len - (uint8_t)dec_varint(d));
    }
    return n;
}
static inline uint16_t dec_huffman_sym(bitdec_t *d, const huff_dec_t *h) { /* coder.rs:144-150 */
    dec_reserve(d);
    uint64_t peeked = bb_peek(&d->bb, h->max_code_len);
    d->bb.len = (uint8_t)(d->bb.len - h->len[peeked]);
    return h->sym[peeked];
}

/* coder.rs:219-266 (test_coder_with_huffman) restated */
int orz_oracle_coder_selftest(const uint8_t *msg, size_t n, uint8_t *enc, size_t enc_cap, size_t *enc_len) {
    uint32_t weights[256] = {0};
    uint8_t lens[256];
    uint16_t codes[256];
    for (size_t i = 0; i < n; i++) weights[msg[i]]++;
    int maxl = orz_oracle_huffman_lens(weights, 256, 15, lens);
    (void)maxl;
    orz_oracle_huffman_codes(lens, 256, codes);
    if (enc_cap < n * 2 + 2048) return -2;
    bitenc_t e = {enc, 0, {0, 0}};
    enc_varint(&e, (uint32_t)n);
    enc_huffman_table(&e, lens, 256);
    for (size_t i = 0; i < n; i++) enc_raw_bits(&e, codes[msg[i]], lens[msg[i]]);
    *enc_len = enc_finish(&e);

    bitdec_t d = {enc, 0, {0, 0}};
    uint32_t num = dec_varint(&d);
    if (num != n) return -1;
    uint8_t dlens[256];
    int dmax;
    int dn = dec_huffman_table(&d, dlens, 256, &dmax);
    if (dn < 0) return -1;
    huff_dec_t *h = (huff_dec_t *)malloc(sizeof *h);
    huff_dec_build(h, dlens, dn, dmax);
    int ok = 0;
    for (size_t i = 0; i < n; i++)
        if (dec_huffman_sym(&d, h) != msg[i]) ok = -1;
    free(h);
    return ok;
}

/* ------------------------------------------------------------------------------------------
 * lz.rs — chunk encoder / decoder
 * ---------------------------------------------------------------------------------------- */
static inline int is_alnum(uint8_t c) { return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z'); }
/* lz.rs:481-486 */
static inline size_t hash1(const uint8_t *buf, size_t pos) { return (size_t)(buf[pos] & 0x7f) | (size_t)is_alnum(buf[pos - 1]) << 7; }
/* lz.rs:488-492 */
static inline size_t hash2(const uint8_t *buf, size_t pos) { return (size_t)(buf[pos] & 0x7f) | hash1(buf, pos - 1) << 7; }

/* lz.rs:494-534 */
static uint8_t g_roid_enc_id[ORZ_BUCKET_ITEMS], g_roid_enc_bitlen[ORZ_BUCKET_ITEMS];
static uint16_t g_roid_enc_bits[ORZ_BUCKET_ITEMS];
static uint16_t g_roid_dec_base[ORZ_ROID_SIZE];
static uint8_t g_roid_dec_bitlen[ORZ_ROID_SIZE];
static int g_roid_ready = 0;
static void roid_tables_init(void) {
    if (g_roid_ready) return;
    size_t base = 0, id = 0, idx = 0;
    while (base < ORZ_BUCKET_ITEMS) { /* lz.rs:494-514 */
        size_t bit_len = id / 2;
        for (size_t rest = 0; rest != ((size_t)1 << bit_len); rest++) {
            if (base < ORZ_BUCKET_ITEMS) {
                g_roid_enc_id[idx] = (uint8_t)id;
                g_roid_enc_bitlen[idx] = (uint8_t)bit_len;
                g_roid_enc_bits[idx] = (uint16_t)rest;
                idx++;
                base++;
            }
        }
        id++;
    }
    base = 0, id = 0, idx = 0;
    while (base < ORZ_BUCKET_ITEMS) { /* lz.rs:516-530 */
        size_t bit_len = id / 2;
        g_roid_dec_base[idx] = (uint16_t)base;
        g_roid_dec_bitlen[idx] = (uint8_t)bit_len;
        idx++;
        id++;
        base += (size_t)1 << bit_len;
    }
    g_roid_ready = 1;
}

/* lz.rs:49-67 */
typedef struct {
    bucket_t *buckets;   /* 256 */
    symrank_t *symranks; /* 512 */
    uint8_t (*words)[2]; /* 32768 */
    int first_block;
    int after_literal;
} lzctx_t;

static int lzctx_init(lzctx_t *c) {
    c->buckets = (bucket_t *)malloc(sizeof(bucket_t) * 256);
    c->symranks = (symrank_t *)malloc(sizeof(symrank_t) * 512);
    c->words = (uint8_t(*)[2])calloc(32768, 2);
    if (!c->buckets || !c->symranks || !c->words) return -1;
    for (int i = 0; i < 256; i++) bucket_init(&c->buckets[i]);
    for (int i = 0; i < 512; i++) symrank_new(&c->symranks[i]);
    c->first_block = 1;
    c->after_literal = 1;
    return 0;
}
static void lzctx_free(lzctx_t *c) {
    free(c->buckets);
    free(c->symranks);
    free(c->words);
}

struct orz_oracle_enc {
    lzctx_t ctx;
    matcher_t *matchers; /* 256 */
    orz_oracle_item *items, *items_pre;
    size_t n_items;
};

orz_oracle_enc *orz_oracle_enc_new(void) { /* lz.rs:75-80 */
    roid_tables_init();
    orz_oracle_enc *e = (orz_oracle_enc *)calloc(1, sizeof *e);
    if (!e) return NULL;
    if (lzctx_init(&e->ctx)) return NULL;
    e->matchers = (matcher_t *)malloc(sizeof(matcher_t) * 256);
    e->items = (orz_oracle_item *)malloc(sizeof(orz_oracle_item) * ORZ_LZ_CHUNK_SIZE);
    e->items_pre = (orz_oracle_item *)malloc(sizeof(orz_oracle_item) * ORZ_LZ_CHUNK_SIZE);
    for (int i = 0; i < 256; i++) matcher_init(&e->matchers[i]);
    return e;
}
void orz_oracle_enc_free(orz_oracle_enc *e) {
    if (!e) return;
    lzctx_free(&e->ctx);
    free(e->matchers);
    free(e->items);
    free(e->items_pre);
    free(e);
}
void orz_oracle_enc_forward(orz_oracle_enc *e, size_t forward_len) { /* lz.rs:82-87 */
    for (int i = 0; i < 256; i++) {
        bucket_forward(&e->ctx.buckets[i], forward_len);
        matcher_forward(&e->matchers[i], &e->ctx.buckets[i]);
    }
}
size_t orz_oracle_enc_last_items(const orz_oracle_enc *e, const orz_oracle_item **pre, const orz_oracle_item **post) {
    if (pre) *pre = e->items_pre;
    if (post) *post = e->items;
    return e->n_items;
}

/* lz.rs:89-345 */
void orz_oracle_enc_encode(orz_oracle_enc *self, const orz_oracle_cfg *cfg, const uint8_t *sbuf, size_t sbuf_len,
                           uint8_t *tbuf, size_t spos, size_t *spos_out, size_t *tlen_out) {
    lzctx_t *ctx = &self->ctx;
    orz_oracle_item *items = self->items;
    size_t n_items = 0;
    bitenc_t enc = {tbuf, 0, {0, 0}};

    /* stage A — lz.rs:131-235 */
    while (spos < sbuf_len && n_items < ORZ_LZ_CHUNK_SIZE) {
        const uint8_t *lwe = ctx->words[hash2(sbuf, spos - 1)];
        uint8_t lwe0 = lwe[0], lwe1 = lwe[1];
        int last_word_matched = sbuf[spos] == lwe0 && sbuf[spos + 1] == lwe1;
        size_t c = hash1(sbuf, spos - 1);
        uint16_t symrank_context = (uint16_t)(c | (size_t)ctx->after_literal << 8);
        uint8_t symrank_unlikely = lwe0;
        int lazy_match_id = 0;
        STAT(g_stats.steps++);

        match_t m = find_match(&self->matchers[c], &ctx->buckets[c], sbuf, sbuf_len, spos, cfg->match_depth);
        if (m.match_len > 0) {
            uint8_t roid = g_roid_enc_id[m.reduced_offset];
            uint8_t robitlen = g_roid_enc_bitlen[m.reduced_offset];
            uint16_t robits = g_roid_enc_bits[m.reduced_offset];

            if (m.match_len < ORZ_LZ_MATCH_MAX_LEN / 2) { /* lz.rs:151-171 */
                size_t lazy_len1 = m.match_len + 1 + (robitlen < 8);
                size_t lazy_len2 = lazy_len1 - (size_t)last_word_matched;
                size_t c1 = hash1(sbuf, spos);
                STAT(g_stats.lazy1_calls++);
                if (has_lazy_match(&self->matchers[c1], &ctx->buckets[c1], sbuf, spos + 1, lazy_len1, cfg->lazy_match_depth1,
                                   &g_stats.lazy1_hops)) {
                    lazy_match_id = 1;
                    STAT(g_stats.lazy1_hits++);
                } else {
                    size_t c2 = hash1(sbuf, spos + 1);
                    STAT(g_stats.lazy2_calls++);
                    if (has_lazy_match(&self->matchers[c2], &ctx->buckets[c2], sbuf, spos + 2, lazy_len2,
                                       cfg->lazy_match_depth2, &g_stats.lazy2_hops)) {
                        lazy_match_id = 2;
                        STAT(g_stats.lazy2_hits++);
                    }
                }
            }

            if (lazy_match_id == 0) { /* lz.rs:173-209 */
                size_t eml_wide;
                if (m.match_len > m.match_len_expected) eml_wide = m.match_len - m.match_len_min;
                else if (m.match_len < m.match_len_expected) eml_wide = m.match_len - m.match_len_min + 1;
                else eml_wide = 0;
                uint8_t eml = (uint8_t)eml_wide; /* `as u8` of wrapping usize arithmetic */
                uint8_t lenid = eml < ORZ_LENID_SIZE - 1 ? eml : ORZ_LENID_SIZE - 1;
                orz_oracle_item *it = &items[n_items++];
                it->symbol = (uint16_t)(256 + (uint16_t)roid * ORZ_LENID_SIZE + lenid);
                it->ctx = symrank_context;
                it->unlikely = symrank_unlikely;
                it->robitlen = robitlen;
                it->robits = robits;
                it->eml = eml;
                it->after_literal = (uint8_t)ctx->after_literal;
                it->is_match = 1;
                it->pad = 0;

                bucket_update(&ctx->buckets[c], spos, m.reduced_offset, m.match_len);
                matcher_update(&self->matchers[c], &ctx->buckets[c], sbuf, spos);
                spos += m.match_len;
                ctx->after_literal = 0;
                size_t w = hash2(sbuf, spos - 3);
                ctx->words[w][0] = sbuf[spos - 2];
                ctx->words[w][1] = sbuf[spos - 1];
                STAT(g_stats.matches++; g_stats.match_len_sum += m.match_len);
                continue;
            }
        }
        bucket_update(&ctx->buckets[c], spos, 0, 0); /* lz.rs:210-215 */
        matcher_update(&self->matchers[c], &ctx->buckets[c], sbuf, spos);

        orz_oracle_item *it = &items[n_items++];
        it->ctx = symrank_context;
        it->unlikely = symrank_unlikely;
        it->robitlen = 0;
        it->robits = 0;
        it->eml = 0;
        it->after_literal = (uint8_t)ctx->after_literal;
        it->is_match = 0;
        it->pad = 0;
        if (spos + 1 < sbuf_len && lazy_match_id != 1 && last_word_matched) { /* lz.rs:218-226 */
            it->symbol = ORZ_WORD_SYMBOL;
            spos += 2;
            ctx->after_literal = 0;
            STAT(g_stats.words++);
        } else { /* lz.rs:227-236 */
            it->symbol = sbuf[spos];
            spos += 1;
            ctx->after_literal = 1;
            size_t w = hash2(sbuf, spos - 3);
            ctx->words[w][0] = sbuf[spos - 2];
            ctx->words[w][1] = sbuf[spos - 1];
            STAT(g_stats.literals++);
        }
    }
    self->n_items = n_items;
    memcpy(self->items_pre, items, sizeof(orz_oracle_item) * n_items);

    /* first chunk of the stream: symbol order header + symrank init — lz.rs:238-265 */
    if (ctx->first_block) {
        uint32_t counts[ORZ_NUM_SYMBOLS] = {0};
        for (size_t k = 0; k < n_items; k++) counts[items[k].symbol]++;
        uint32_t num_counted = 0;
        for (int s = 0; s < ORZ_NUM_SYMBOLS; s++) num_counted += counts[s] > 1;
        /* stable sort by Reverse(max(count,1)): insertion sort keeps equal keys in symbol order */
        uint16_t vs[ORZ_NUM_SYMBOLS];
        for (int s = 0; s < ORZ_NUM_SYMBOLS; s++) {
            uint32_t key = counts[s] > 1 ? counts[s] : 1;
            int j = s;
            while (j > 0) {
                uint32_t kj = counts[vs[j - 1]] > 1 ? counts[vs[j - 1]] : 1;
                if (kj >= key) break;
                vs[j] = vs[j - 1];
                j--;
            }
            vs[j] = (uint16_t)s;
        }
        enc_varint(&enc, num_counted);
        for (uint32_t k = 0; k < num_counted; k++) enc_raw_bits(&enc, vs[k], 9);
        symrank_t init;
        symrank_new(&init);
        symrank_init(&init, vs);
        for (int i = 0; i < 512; i++) ctx->symranks[i] = init;
        ctx->first_block = 0;
    }

    /* lz.rs:268-269 */
    enc_varint(&enc, (uint32_t)(spos < sbuf_len ? spos : sbuf_len));
    enc_varint(&enc, (uint32_t)n_items);

    /* stage B — lz.rs:272-305 */
    uint32_t w1[2][ORZ_NUM_SYMBOLS];
    uint32_t w2[ORZ_LZ_MATCH_MAX_LEN];
    memset(w1, 0, sizeof w1);
    memset(w2, 0, sizeof w2);
    for (size_t k = 0; k < n_items; k++) {
        orz_oracle_item *it = &items[k];
        uint16_t r = symrank_encode(&ctx->symranks[it->ctx], it->symbol, it->unlikely);
        w1[it->after_literal][r]++;
        if (it->is_match && it->eml >= ORZ_LENID_SIZE - 1) w2[it->eml]++;
        it->symbol = r;
    }

    /* stage C — lz.rs:306-318 */
    uint8_t l1[2][ORZ_NUM_SYMBOLS], l2[ORZ_LZ_MATCH_MAX_LEN];
    uint16_t c1[2][ORZ_NUM_SYMBOLS], c2[ORZ_LZ_MATCH_MAX_LEN];
    orz_oracle_huffman_lens(w1[0], ORZ_NUM_SYMBOLS, 15, l1[0]);
    orz_oracle_huffman_lens(w1[1], ORZ_NUM_SYMBOLS, 15, l1[1]);
    orz_oracle_huffman_lens(w2, ORZ_LZ_MATCH_MAX_LEN, 15, l2);
    enc_huffman_table(&enc, l1[0], ORZ_NUM_SYMBOLS);
    enc_huffman_table(&enc, l1[1], ORZ_NUM_SYMBOLS);
    enc_huffman_table(&enc, l2, ORZ_LZ_MATCH_MAX_LEN);
    orz_oracle_huffman_codes(l1[0], ORZ_NUM_SYMBOLS, c1[0]);
    orz_oracle_huffman_codes(l1[1], ORZ_NUM_SYMBOLS, c1[1]);
    orz_oracle_huffman_codes(l2, ORZ_LZ_MATCH_MAX_LEN, c2);

    /* stage D — lz.rs:320-342 */
    for (size_t k = 0; k < n_items; k++) {
        const orz_oracle_item *it = &items[k];
        enc_raw_bits(&enc, c1[it->after_literal][it->symbol], l1[it->after_literal][it->symbol]);
        if (it->is_match) {
            enc_raw_bits(&enc, it->robits, it->robitlen);
            if (it->eml >= ORZ_LENID_SIZE - 1) enc_raw_bits(&enc, c2[it->eml], l2[it->eml]);
        }
    }
    *spos_out = spos;
    *tlen_out = enc_finish(&enc); /* lz.rs:344 */
}

struct orz_oracle_dec {
    lzctx_t ctx;
    huff_dec_t *h1[2], *h2;
};

orz_oracle_dec *orz_oracle_dec_new(void) { /* lz.rs:353-357 */
    roid_tables_init();
    orz_oracle_dec *d = (orz_oracle_dec *)calloc(1, sizeof *d);
    if (!d) return NULL;
    if (lzctx_init(&d->ctx)) return NULL;
    d->h1[0] = (huff_dec_t *)malloc(sizeof(huff_dec_t));
    d->h1[1] = (huff_dec_t *)malloc(sizeof(huff_dec_t));
    d->h2 = (huff_dec_t *)malloc(sizeof(huff_dec_t));
    return d;
}
void orz_oracle_dec_free(orz_oracle_dec *d) {
    if (!d) return;
    lzctx_free(&d->ctx);
    free(d->h1[0]);
    free(d->h1[1]);
    free(d->h2);
    free(d);
}
void orz_oracle_dec_forward(orz_oracle_dec *d, size_t forward_len) { /* lz.rs:359-364 */
    for (int i = 0; i < 256; i++) bucket_forward(&d->ctx.buckets[i], forward_len);
}

/* test hook: per decoded item (position, value | match length << 16, reduced offset) of the chunks decoded while it is set */
static uint32_t *g_dec_trace;
static size_t g_dec_trace_cap, g_dec_trace_n;
void orz_oracle_dec_set_trace(uint32_t *buf, size_t cap_items) {
    g_dec_trace = buf;
    g_dec_trace_cap = cap_items;
    g_dec_trace_n = 0;
}
size_t orz_oracle_dec_trace_len(void) { return g_dec_trace_n; }
static inline void dec_trace(size_t spos, uint32_t v, size_t match_len, size_t reduced_offset) {
    if (g_dec_trace && g_dec_trace_n < g_dec_trace_cap) {
        uint32_t *t = g_dec_trace + 3 * g_dec_trace_n++;
        t[0] = (uint32_t)spos;
        t[1] = v | (uint32_t)match_len << 16;
        t[2] = (uint32_t)reduced_offset;
    }
}

/* lz.rs:366-478.  tbuf must be readable for 8 bytes past tlen (decoder refill over-read). */
int orz_oracle_dec_decode(orz_oracle_dec *self, const uint8_t *tbuf, size_t tlen, uint8_t *sbuf, size_t spos, size_t *spos_end) {
    lzctx_t *ctx = &self->ctx;
    bitdec_t dec = {tbuf, 0, {0, 0}};
    (void)tlen;

    if (ctx->first_block) { /* lz.rs:372-392 */
        size_t num = dec_varint(&dec);
        uint16_t vs[ORZ_NUM_SYMBOLS] = {0};
        uint8_t set[ORZ_NUM_SYMBOLS] = {0};
        if (num > ORZ_NUM_SYMBOLS) return -1; /* the reference would index out of bounds (UB); we refuse */
        for (size_t i = 0; i < num; i++) {
            vs[i] = (uint16_t)dec_raw_bits(&dec, 9);
            if (vs[i] >= ORZ_NUM_SYMBOLS) return -1;
            set[vs[i]] = 1;
        }
        for (int i = 0; i < ORZ_NUM_SYMBOLS; i++)
            if (!set[i]) {
                if (num >= ORZ_NUM_SYMBOLS) return -1;
                vs[num++] = (uint16_t)i;
            }
        symrank_t init;
        symrank_new(&init);
        symrank_init(&init, vs);
        for (int i = 0; i < 512; i++) ctx->symranks[i] = init;
        ctx->first_block = 0;
    }

    size_t sbuf_len = dec_varint(&dec); /* lz.rs:396-397 */
    size_t n_items = dec_varint(&dec);

    uint8_t l1[2][1024], l2[1024];
    int m1[2], m2, n1[2], n2;
    n1[0] = dec_huffman_table(&dec, l1[0], 1024, &m1[0]); /* lz.rs:400-409 */
    n1[1] = dec_huffman_table(&dec, l1[1], 1024, &m1[1]);
    n2 = dec_huffman_table(&dec, l2, 1024, &m2);
    if (n1[0] < 0 || n1[1] < 0 || n2 < 0 || m1[0] > 16 || m1[1] > 16 || m2 > 16) return -1;
    huff_dec_build(self->h1[0], l1[0], n1[0], m1[0]);
    huff_dec_build(self->h1[1], l1[1], n1[1], m1[1]);
    huff_dec_build(self->h2, l2, n2, m2);

    for (size_t k = 0; k < n_items; k++) { /* lz.rs:411-476 */
        uint16_t symbol = dec_huffman_sym(&dec, self->h1[ctx->after_literal]);
        if (symbol > ORZ_NUM_SYMBOLS) return -1; /* lz.rs:413-415 */
        if (symbol == ORZ_NUM_SYMBOLS) return -1; /* the reference lets 389 through into an unchecked index (UB) */

        bucket_t *cur = &ctx->buckets[hash1(sbuf, spos - 1)];
        size_t wi = hash2(sbuf, spos - 1);
        uint8_t lwe0 = ctx->words[wi][0], lwe1 = ctx->words[wi][1];
        uint16_t sctx = (uint16_t)(hash1(sbuf, spos - 1) | (size_t)ctx->after_literal << 8);
        uint16_t v = symrank_decode(&ctx->symranks[sctx], symbol, lwe0);

        if (v == ORZ_WORD_SYMBOL) { /* lz.rs:425-430 */
            dec_trace(spos, v, 0, 0);
            bucket_update(cur, spos, 0, 0);
            ctx->after_literal = 0;
            sbuf[spos] = lwe0;
            sbuf[spos + 1] = lwe1;
            spos += 2;
        } else if (v < 256) { /* lz.rs:431-437 */
            dec_trace(spos, v, 0, 0);
            bucket_update(cur, spos, 0, 0);
            ctx->after_literal = 1;
            sbuf[spos] = (uint8_t)v;
            spos += 1;
            size_t w = hash2(sbuf, spos - 3);
            ctx->words[w][0] = sbuf[spos - 2];
            ctx->words[w][1] = sbuf[spos - 1];
        } else { /* lz.rs:438-474 */
            uint8_t roid = (uint8_t)((v - 256) / ORZ_LENID_SIZE);
            uint8_t lenid = (uint8_t)((v - 256) % ORZ_LENID_SIZE);
            size_t reduced_offset = (size_t)g_roid_dec_base[roid] + dec_raw_bits(&dec, g_roid_dec_bitlen[roid]);
            uint32_t node = node_sub(cur->head, (uint32_t)reduced_offset);
            uint32_t n1v = cur->nodes1[node];
            size_t match_len_expected = cur->nodes2[node];
            size_t eml = lenid == ORZ_LENID_SIZE - 1 ? dec_huffman_sym(&dec, self->h2) : lenid;
            size_t match_pos = NODE_POS(n1v);
            size_t match_len_min = NODE_LEN_MIN(n1v);
            if (match_len_min < ORZ_LZ_MATCH_MIN_LEN) match_len_min = ORZ_LZ_MATCH_MIN_LEN;
            if (match_len_expected < ORZ_LZ_MATCH_MIN_LEN) match_len_expected = ORZ_LZ_MATCH_MIN_LEN;
            size_t match_len;
            if (eml + match_len_min > match_len_expected) match_len = eml + match_len_min;
            else if (eml > 0) match_len = eml + match_len_min - 1;
            else match_len = match_len_expected;
            dec_trace(spos, v, match_len, reduced_offset);
            bucket_update(cur, spos, reduced_offset, match_len);
            ctx->after_literal = 0;
            if (match_pos >= spos) return -1; /* corrupt stream: the reference would loop forever / UB */
            mem_fast_copy(sbuf, match_pos, spos, match_len);
            spos += match_len;
            size_t w = hash2(sbuf, spos - 3);
            ctx->words[w][0] = sbuf[spos - 2];
            ctx->words[w][1] = sbuf[spos - 1];
        }
        if (spos > ORZ_LZ_BLOCK_SIZE + ORZ_LZ_MATCH_MAX_LEN) return -1; /* corrupt stream guard */
    }
    *spos_end = spos < sbuf_len ? spos : sbuf_len; /* lz.rs:477 */
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * main.rs:97-102 — levels
 * ---------------------------------------------------------------------------------------- */
int orz_oracle_level_cfg(int level, orz_oracle_cfg *cfg) {
    switch (level) {
    case 0: cfg->match_depth = 5, cfg->lazy_match_depth1 = 3, cfg->lazy_match_depth2 = 2; return 0;
    case 1: cfg->match_depth = 15, cfg->lazy_match_depth1 = 9, cfg->lazy_match_depth2 = 6; return 0;
    case 2: cfg->match_depth = 45, cfg->lazy_match_depth1 = 27, cfg->lazy_match_depth2 = 18; return 0;
    default: return -1;
    }
}

/* ------------------------------------------------------------------------------------------
 * lib.rs:58-129 + ioutil.rs:60-88 — stream driver, memory to memory
 * ---------------------------------------------------------------------------------------- */
size_t orz_oracle_encode_bound(size_t n) {
    /* generous: payload <= ~40 bits/item, >= 1 byte/item; plus per-chunk tables and framing */
    return n + n / 2 + (n / 65536 + 2) * 4096 + 65536;
}

static size_t write_len(uint8_t *out, size_t len) { /* ioutil.rs:79-88 */
    size_t k = 0;
    while (len >= 128) {
        out[k++] = (uint8_t)(128 + len % 128);
        len /= 128;
    }
    out[k++] = (uint8_t)len;
    return k;
}

int orz_oracle_encode_stream(const orz_oracle_cfg *cfg, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap, size_t *out_len) {
    orz_oracle_enc *lzenc = orz_oracle_enc_new(); /* lib.rs:64 */
    uint8_t *sbvec_buf = (uint8_t *)calloc(ORZ_LZ_BLOCK_SIZE + ORZ_SENTINEL_LEN * 2, 1); /* lib.rs:67 */
    uint8_t *tbvec = (uint8_t *)calloc(ORZ_TBUF_LEN, 1);                               /* lib.rs:68 */
    if (!lzenc || !sbvec_buf || !tbvec) return -4;
    uint8_t *sbvec = sbvec_buf + ORZ_SENTINEL_LEN; /* lib.rs:69 */
    size_t ipos = 0, opos = 0;
    int rc = 0;

    for (;;) { /* lib.rs:72-86 */
        size_t room = ORZ_LZ_BLOCK_SIZE - ORZ_PREMATCH_LEN;
        size_t rd = n - ipos < room ? n - ipos : room;
        if (rd == 0) break;
        memcpy(sbvec + ORZ_PREMATCH_LEN, in + ipos, rd);
        ipos += rd;
        size_t spos = ORZ_PREMATCH_LEN;
        while (spos < ORZ_PREMATCH_LEN + rd) {
            size_t s, t;
            orz_oracle_enc_encode(lzenc, cfg, sbvec, ORZ_PREMATCH_LEN + rd, tbvec, spos, &s, &t);
            if (opos + 5 + t + 1 > out_cap) {
                rc = -2;
                goto done;
            }
            opos += write_len(out + opos, t);
            memcpy(out + opos, tbvec, t);
            opos += t;
            spos = s;
        }
        memmove(sbvec, sbvec + ORZ_LZ_BLOCK_SIZE - ORZ_PREMATCH_LEN, ORZ_PREMATCH_LEN); /* lib.rs:83 */
        orz_oracle_enc_forward(lzenc, ORZ_LZ_BLOCK_SIZE - ORZ_PREMATCH_LEN);           /* lib.rs:84 */
    }
    if (opos + 1 > out_cap) {
        rc = -2;
        goto done;
    }
    out[opos++] = 0; /* lib.rs:89 */
    *out_len = opos;
done:
    orz_oracle_enc_free(lzenc);
    free(sbvec_buf);
    free(tbvec);
    return rc;
}

int orz_oracle_decode_stream(const uint8_t *in, size_t n, uint8_t *out, size_t out_cap, size_t *out_len) {
    orz_oracle_dec *lzdec = orz_oracle_dec_new(); /* lib.rs:99 */
    uint8_t *sbvec_buf = (uint8_t *)calloc((size_t)ORZ_LZ_BLOCK_SIZE * 2 + ORZ_SENTINEL_LEN * 2, 1); /* lib.rs:102 */
    uint8_t *tbvec = (uint8_t *)calloc(ORZ_TBUF_LEN + 16, 1);                                     /* lib.rs:103 */
    if (!lzdec || !sbvec_buf || !tbvec) return -4;
    uint8_t *sbvec = sbvec_buf + ORZ_SENTINEL_LEN;
    size_t ipos = 0, opos = 0, spos = ORZ_PREMATCH_LEN;
    int rc = 0;

    for (;;) { /* lib.rs:107-126 */
        size_t t = 0, factor = 1; /* ioutil.rs:60-76 */
        for (;;) {
            if (ipos >= n) {
                rc = -3;
                goto done;
            }
            uint8_t v = in[ipos++];
            if (v < 128) {
                t += (size_t)v * factor;
                break;
            }
            t += (size_t)(v - 128) * factor;
            factor *= 128;
        }
        if (t == 0) break;
        if (t >= ORZ_TBUF_LEN) { /* lib.rs:111-113 */
            rc = -1;
            goto done;
        }
        if (ipos + t > n) { /* lib.rs:114 read_exact */
            rc = -3;
            goto done;
        }
        memcpy(tbvec, in + ipos, t);
        ipos += t;
        size_t spos_end;
        if (orz_oracle_dec_decode(lzdec, tbvec, t, sbvec, spos, &spos_end)) {
            rc = -1;
            goto done;
        }
        if (spos_end < spos || spos_end > ORZ_LZ_BLOCK_SIZE) {
            rc = -1;
            goto done;
        }
        if (opos + (spos_end - spos) > out_cap) {
            rc = -2;
            goto done;
        }
        memcpy(out + opos, sbvec + spos, spos_end - spos); /* lib.rs:117 */
        opos += spos_end - spos;
        spos = spos_end;
        if (spos >= ORZ_LZ_BLOCK_SIZE) { /* lib.rs:120-125 */
            memmove(sbvec, sbvec + ORZ_LZ_BLOCK_SIZE - ORZ_PREMATCH_LEN, ORZ_PREMATCH_LEN);
            orz_oracle_dec_forward(lzdec, ORZ_LZ_BLOCK_SIZE - ORZ_PREMATCH_LEN);
            spos = ORZ_PREMATCH_LEN;
        }
    }
    *out_len = opos;
done:
    orz_oracle_dec_free(lzdec);
    free(sbvec_buf);
    free(tbvec);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * optional CLI: orz_oracle encode -l N in out | decode in out   (build with -DORZ_ORACLE_MAIN)
 * ---------------------------------------------------------------------------------------- */
#ifdef ORZ_ORACLE_MAIN
#include <stdio.h>
#include <time.h>
static uint8_t *read_file(const char *path, size_t *n) {
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *b = (uint8_t *)malloc((size_t)sz + 1);
    *n = fread(b, 1, (size_t)sz, f);
    fclose(f);
    return b;
}
static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}
int main(int argc, char **argv) {
    if (argc >= 5 && !strcmp(argv[1], "encode")) {
        int level = atoi(argv[2]);
        orz_oracle_cfg cfg;
        if (orz_oracle_level_cfg(level, &cfg)) return fprintf(stderr, "invalid level: %d\n", level), 1;
        size_t n, on;
        uint8_t *in = read_file(argv[3], &n);
        if (!in) return fprintf(stderr, "cannot read %s\n", argv[3]), 1;
        size_t cap = orz_oracle_encode_bound(n);
        uint8_t *out = (uint8_t *)malloc(cap);
        if (getenv("ORZ_STATS")) orz_oracle_stats_reset();
        double t0 = now_s();
        int rc = orz_oracle_encode_stream(&cfg, in, n, out, cap, &on);
        double t1 = now_s();
        if (rc) return fprintf(stderr, "encoding failed: %d\n", rc), 1;
        FILE *f = fopen(argv[4], "wb");
        fwrite(out, 1, on, f);
        fclose(f);
        fprintf(stderr, "encode: %zu => %zu bytes, %.3f s, %.2f MB/s\n", n, on, t1 - t0, n / (t1 - t0) * 1e-6);
        if (getenv("ORZ_STATS")) {
            orz_oracle_stats s;
            orz_oracle_stats_get(&s);
            fprintf(stderr,
                    "steps %llu matches %llu words %llu literals %llu avg_match_len %.2f\n"
                    "find_match: calls %llu hops %llu (%.2f/call) probe_hits %llu\n"
                    "lazy1: calls %llu hops %llu hits %llu | lazy2: calls %llu hops %llu hits %llu\n",
                    (unsigned long long)s.steps, (unsigned long long)s.matches, (unsigned long long)s.words,
                    (unsigned long long)s.literals, s.matches ? (double)s.match_len_sum / s.matches : 0.0,
                    (unsigned long long)s.fm_calls, (unsigned long long)s.fm_hops, s.fm_calls ? (double)s.fm_hops / s.fm_calls : 0.0,
                    (unsigned long long)s.fm_probe_hits, (unsigned long long)s.lazy1_calls, (unsigned long long)s.lazy1_hops,
                    (unsigned long long)s.lazy1_hits, (unsigned long long)s.lazy2_calls, (unsigned long long)s.lazy2_hops,
                    (unsigned long long)s.lazy2_hits);
        }
        return 0;
    }
    if (argc >= 4 && !strcmp(argv[1], "decode")) {
        size_t n, on;
        uint8_t *in = read_file(argv[2], &n);
        if (!in) return fprintf(stderr, "cannot read %s\n", argv[2]), 1;
        size_t cap = n * 12 + (1 << 20);
        uint8_t *out = (uint8_t *)malloc(cap);
        double t0 = now_s();
        int rc = orz_oracle_decode_stream(in, n, out, cap, &on);
        double t1 = now_s();
        if (rc) return fprintf(stderr, "decoding failed: %d\n", rc), 1;
        FILE *f = fopen(argv[3], "wb");
        fwrite(out, 1, on, f);
        fclose(f);
        fprintf(stderr, "decode: %zu => %zu bytes, %.3f s, %.2f MB/s\n", n, on, t1 - t0, on / (t1 - t0) * 1e-6);
        return 0;
    }
    fprintf(stderr, "usage: %s encode LEVEL in out | decode in out\n", argv[0]);
    return 2;
}
#endif
