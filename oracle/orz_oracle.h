/*
 * orz_oracle.h — CPU restatement of richox/orz's ROLZ stream codec (TEST INFRASTRUCTURE).
 *
 * This is the parity oracle for the orz_b200 CUDA path.  It is NOT part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may link or execute it.  Every function cites the reference file:line (relative to the
 * richox/orz tree @29414c6) that it restates.
 *
 * Pinning: PARITY UNPINNED against real `orz` bytes — the reference ships no golden bitstream vectors (its tests are
 * round-trips only) and cannot be built in this environment.  The oracle is pinned by the re-expressed reference unit
 * test, hand-derived vectors, round-trips, and enwik8 sizes that are within 0.13 % of the README's (README.md:43-46),
 * consistently smaller at all three levels; see orz_oracle.c and DESIGN.md §1.
 */
#ifndef ORZ_ORACLE_H
#define ORZ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/lib.rs:31-34, 54-55 ; src/lz.rs:24-29 ; src/matcher.rs:18 */
#ifdef ORZ_TEST_BLOCK_BITS /* scale model for the emulator tests (tests/simt/Makefile); never set for liborz_oracle.so */
#define ORZ_LZ_BLOCK_SIZE        ((1u << ORZ_TEST_BLOCK_BITS) - 1)
#define ORZ_LZ_CHUNK_SIZE        (ORZ_TEST_CHUNK_ITEMS)
#else
#define ORZ_LZ_BLOCK_SIZE        ((1u << 25) - 1)
#define ORZ_LZ_CHUNK_SIZE        (1u << 20)
#endif
#define ORZ_LZ_MATCH_MAX_LEN     240
#define ORZ_LZ_MATCH_MIN_LEN     4
#define ORZ_SENTINEL_LEN         (ORZ_LZ_MATCH_MAX_LEN * 2)
#define ORZ_PREMATCH_LEN         (ORZ_LZ_BLOCK_SIZE / 2)
#define ORZ_BUCKET_ITEMS         4094
#define ORZ_BUCKET_HASH          4627 /* (4094*1.13) as usize | 1 */
#define ORZ_ROID_SIZE            22
#define ORZ_LENID_SIZE           6
#define ORZ_NUM_SYMBOLS          (256 + ORZ_ROID_SIZE * ORZ_LENID_SIZE + 1) /* 389 */
#define ORZ_WORD_SYMBOL          (ORZ_NUM_SYMBOLS - 1)
#define ORZ_TBUF_LEN             (ORZ_PREMATCH_LEN * 3)

/* src/lz.rs:32-37 (#[repr(C)] LZCfg) */
typedef struct {
    size_t match_depth;
    size_t lazy_match_depth1;
    size_t lazy_match_depth2;
} orz_oracle_cfg;

/* One parsed step of stage A (src/lz.rs:100-117, MatchItem). `symbol` holds the pre-rank
 * symbol after stage A and the rank after stage B, like the reference's in-place rewrite. */
typedef struct {
    uint16_t symbol;
    uint16_t ctx;       /* symrank_context, 9 bits */
    uint16_t robits;
    uint8_t  unlikely;  /* symrank_unlikely */
    uint8_t  robitlen;
    uint8_t  eml;       /* encoded_match_len */
    uint8_t  after_literal;
    uint8_t  is_match;
    uint8_t  pad;
} orz_oracle_item;

typedef struct orz_oracle_enc orz_oracle_enc;
typedef struct orz_oracle_dec orz_oracle_dec;

/* main.rs:97-102: level -> LZCfg; returns -1 for an invalid level */
int orz_oracle_level_cfg(int level, orz_oracle_cfg *cfg);

/* ---- chunk-level operator interface (src/lz.rs:75,82,89 ; 353,359,366) ---- */
orz_oracle_enc *orz_oracle_enc_new(void);
void orz_oracle_enc_free(orz_oracle_enc *);
void orz_oracle_enc_forward(orz_oracle_enc *, size_t forward_len);
/* sbuf must have >= ORZ_SENTINEL_LEN readable bytes before sbuf[0] and after sbuf[sbuf_len-1]. */
void orz_oracle_enc_encode(orz_oracle_enc *, const orz_oracle_cfg *, const uint8_t *sbuf, size_t sbuf_len,
                           uint8_t *tbuf, size_t spos, size_t *spos_out, size_t *tlen_out);
/* test hook: items of the last encoded chunk, pre-rank (stage A output) and post-rank (stage B output) */
size_t orz_oracle_enc_last_items(const orz_oracle_enc *, const orz_oracle_item **pre, const orz_oracle_item **post);

orz_oracle_dec *orz_oracle_dec_new(void);
void orz_oracle_dec_free(orz_oracle_dec *);
void orz_oracle_dec_forward(orz_oracle_dec *, size_t forward_len);
/* returns 0, or -1 for InvalidData (lz.rs:413-415) */
int orz_oracle_dec_decode(orz_oracle_dec *, const uint8_t *tbuf, size_t tlen, uint8_t *sbuf, size_t spos, size_t *spos_end);

/* test hook: while set, every decoded item appends (position, value | match length << 16, reduced offset) to buf (3 words per item) */
void orz_oracle_dec_set_trace(uint32_t *buf, size_t cap_items);
size_t orz_oracle_dec_trace_len(void);

/* ---- stream level (src/lib.rs:58-92, 94-129), memory to memory ---- */
size_t orz_oracle_encode_bound(size_t n);
/* returns 0 on success, -2 if out_cap is too small */
int orz_oracle_encode_stream(const orz_oracle_cfg *, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap, size_t *out_len);
/* returns 0, -1 InvalidData, -2 output too small, -3 UnexpectedEof */
int orz_oracle_decode_stream(const uint8_t *in, size_t n, uint8_t *out, size_t out_cap, size_t *out_len);

/* ---- pieces exported for unit tests ---- */
/* huffman.rs:27-111: code lengths (<= max_code_len) from weights; returns max code length */
int orz_oracle_huffman_lens(const uint32_t *weights, int n, int max_code_len, uint8_t *lens);
/* huffman.rs:119-141: canonical codes */
void orz_oracle_huffman_codes(const uint8_t *lens, int n, uint16_t *codes);
/* matcher.rs:255-263 */
uint32_t orz_oracle_hash_dword(const uint8_t *p);
/* coder.rs:219-266 re-expressed: encodes varint(n)+table+symbols, decodes, returns 0 if equal; *enc_len = bytes */
int orz_oracle_coder_selftest(const uint8_t *msg, size_t n, uint8_t *enc, size_t enc_cap, size_t *enc_len);
/* symrank.rs: run encode() over a sequence on a fresh identity-initialised coder (test hook) */
void orz_oracle_symrank_run(const uint16_t *init_order, const uint16_t *syms, const uint16_t *unlikely, size_t n,
                            uint16_t *ranks_out);

/* ---- optional instrumentation (stage-A statistics for kernel design; zero cost unless enabled) ---- */
typedef struct {
    uint64_t steps, matches, words, literals;
    uint64_t fm_calls, fm_hops, fm_probe_hits, fm_lcp_calls;
    uint64_t lazy1_calls, lazy1_hops, lazy2_calls, lazy2_hops;
    uint64_t lazy1_hits, lazy2_hits;
    uint64_t match_len_sum;
} orz_oracle_stats;
void orz_oracle_stats_reset(void);
void orz_oracle_stats_get(orz_oracle_stats *);

#ifdef __cplusplus
}
#endif
#endif
