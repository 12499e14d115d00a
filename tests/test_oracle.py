"""CPU tests of the oracle (oracle/orz_oracle.c): the checker itself must be pinned before it may judge the GPU.

Pins available: the reference's only unit test (coder.rs:219-266) re-expressed, hand-derived vectors
(SURVEY App. D), round trips at every level including chunk cuts and block slides, the committed golden hashes, and
the whole-enwik8 sizes next to the README's (README.md:43-46; they differ by 0.13 %).
"""
import hashlib
import heapq
import json
import os

import numpy as np
import pytest

from oracle import pyoracle
from tests import datagen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))


def test_hand_derived_vectors():
    # SURVEY Appendix D, derived by reading lib.rs/lz.rs/coder.rs (not produced by the reference binary)
    assert pyoracle.encode(b"", 2) == bytes.fromhex("00")
    for level in (0, 1, 2):
        assert pyoracle.encode(b"a", level) == bytes.fromhex("0c2aaaaaaaaaaa941eab40000000")


def test_reference_coder_unit_test():
    # coder.rs:219-266 test_coder_with_huffman: varint + table + symbols round-trip
    import ctypes as C
    msg = b"i can can a can into a can, can you can a can into a can?"
    enc = np.zeros(8192, dtype=np.uint8)
    n = C.c_size_t()
    rc = pyoracle.lib().orz_oracle_coder_selftest(msg, len(msg), enc.ctypes.data, enc.size, C.byref(n))
    assert rc == 0 and 0 < n.value < len(msg) and n.value % 4 == 0  # coder.rs:75-82 pads to 32 bits


@pytest.mark.parametrize("level", [0, 1, 2])
@pytest.mark.parametrize("gen,n", [("text_like", 200_000), ("random_bytes", 50_000), ("low_entropy", 300_000)])
def test_round_trip(gen, n, level):
    data = getattr(datagen, gen)(n)
    enc = pyoracle.encode(data, level)
    assert enc[-1] == 0  # lib.rs:89 EOF marker
    assert pyoracle.decode(enc, n + 64) == data


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 5, 7, 8, 15, 16, 17, 239, 240, 241, 479, 480, 481])
def test_round_trip_tiny_sizes(n):
    data = datagen.text_like(max(n, 1), seed=n)[:n]
    for level in (0, 2):
        assert pyoracle.decode(pyoracle.encode(data, level), n + 64) == data


def test_chunk_cut_and_block_slide_round_trip():
    data = datagen.random_bytes(1_100_000, seed=5)  # > 2^20 items: two chunks (lz.rs:131)
    enc = pyoracle.encode(data, 0)
    assert pyoracle.decode(enc, len(data) + 64) == data
    big = (datagen.text_like(3_000_000, seed=6) * 6)[: (1 << 24) + 12345]  # one full block + a short one (lib.rs:83-84)
    enc = pyoracle.encode(big, 0)
    assert pyoracle.decode(enc, len(big) + 64) == big


def test_payload_framing():
    data = datagen.text_like(100_000, seed=8)
    enc = pyoracle.encode(data, 1)
    # walk the frames: ULEB128(t) payload[t] ... 0x00 (ioutil.rs:60-88, lib.rs:79-80,89)
    pos, frames = 0, 0
    while True:
        t, shift = 0, 1
        while True:
            v = enc[pos]
            pos += 1
            if v < 128:
                t += v * shift
                break
            t += (v - 128) * shift
            shift *= 128
        if t == 0:
            break
        assert t % 4 == 0 and t < pyoracle.TBUF_LEN
        pos += t
        frames += 1
    assert pos == len(enc) and frames == 1


def test_huffman_properties():
    rng = np.random.default_rng(3)
    for trial in range(40):
        n = int(rng.integers(2, 389))
        w = (rng.pareto(0.7, size=n) * 50).astype(np.uint32)
        w[rng.random(n) < 0.3] = 0
        if (w > 0).sum() < 2:
            w[:2] = [3, 5]
        lens = pyoracle.huffman_lens(w, 15)
        used = w > 0
        assert (lens[~used] == 0).all() and (lens[used] >= 1).all() and lens.max() <= 15
        assert abs(sum(2.0 ** -int(l) for l in lens[used]) - 1.0) < 1e-12  # a full binary tree: Kraft equality
        codes = pyoracle.huffman_codes(lens)
        # canonical: ordered by (len, sym), prefix-free (huffman.rs:130-139)
        order = sorted((int(lens[s]), s) for s in range(n) if lens[s])
        vals = [int(codes[s]) << (15 - l) for l, s in order]
        assert vals == sorted(vals) and len(set(vals)) == len(vals)
        # an independent construction with the reference's tie-break (weight, index; internal nodes numbered after the
        # leaves, huffman.rs:28-38,73-88): identical lengths whenever the 15-bit limit is not hit
        h = [(int(x), i) for i, x in enumerate(w) if x]
        heapq.heapify(h)
        kids, nxt = {}, n
        while len(h) > 1:
            a, b = heapq.heappop(h), heapq.heappop(h)
            kids[nxt] = (a[1], b[1])
            heapq.heappush(h, (a[0] + b[0], nxt))
            nxt += 1
        depth = {h[0][1]: 0}
        for node in range(nxt - 1, n - 1, -1):
            for ch in kids[node]:
                depth[ch] = depth[node] + 1
        ref = np.array([depth.get(i, 0) for i in range(n)])
        if ref.max() <= 15:
            assert (lens == ref).all()


def test_huffman_length_limit_and_degenerate():
    fib = [1, 1]
    while len(fib) < 40:
        fib.append(fib[-1] + fib[-2])
    lens = pyoracle.huffman_lens(np.array(fib, dtype=np.uint32), 15)  # unconstrained depth would be 39
    # huffman.rs:99-108 divides every weight by 2^(39-15) and rebuilds: a crude but deterministic limiter
    assert lens.max() <= 15 and (lens > 0).all()
    assert abs(sum(2.0 ** -int(l) for l in lens) - 1.0) < 1e-12
    one = np.zeros(389, dtype=np.uint32)
    one[77] = 9
    lens = pyoracle.huffman_lens(one)
    assert lens[77] == 1 and lens.sum() == 1  # huffman.rs:64-69
    assert pyoracle.huffman_lens(np.zeros(240, dtype=np.uint32)).sum() == 0  # huffman.rs:70


def test_hash_dword_matches_formula():
    rng = np.random.default_rng(9)
    for _ in range(200):
        b = rng.integers(0, 256, size=4, dtype=np.uint8)
        want = sum(((int(x) * m) & 0xFFFFFFFF) ^ a for x, m, a in
                   zip(b, [131313131, 1313131, 13131, 131], [797, 79797, 7979797, 797979797])) & 0xFFFFFFFF
        assert pyoracle.lib().orz_oracle_hash_dword(b.ctypes.data) == want  # matcher.rs:255-263


def test_golden_vectors():
    cases = {"empty": b"", "a": b"a", "abcabc": b"abcabcabcabcabcabcabcabc", "fox": b"the quick brown fox jumps over the lazy dog. " * 20,
             "text_64k": datagen.text_like(65536, seed=101), "random_20k": datagen.random_bytes(20000, seed=102),
             "lowent_100k": datagen.low_entropy(100000, seed=103)}
    for name, data in cases.items():
        g = GOLDEN["cases"][name]
        assert hashlib.sha256(data).hexdigest() == g["sha256_in"], f"generator drifted: {name}"
        for level in ("0", "1", "2"):
            enc = pyoracle.encode(data, int(level))
            assert len(enc) == g["levels"][level]["len"]
            assert hashlib.sha256(enc).hexdigest() == g["levels"][level]["sha256"]
            if "hex" in g["levels"][level]:
                assert enc.hex() == g["levels"][level]["hex"]


def test_enwik8_prefix_golden():
    e8 = datagen.enwik8(2_000_000)
    if e8 is None:
        pytest.skip("data/enwik8.xz not present")
    g = GOLDEN["enwik8_2M"]
    assert hashlib.sha256(e8).hexdigest() == g["sha256_in"]
    for level in ("0", "2"):
        enc = pyoracle.encode(e8, int(level))
        assert (len(enc), hashlib.sha256(enc).hexdigest()) == (g["levels"][level]["len"], g["levels"][level]["sha256"])


def test_decode_rejects_garbage():
    data = datagen.text_like(20_000, seed=12)
    enc = bytearray(pyoracle.encode(data, 1))
    with pytest.raises(RuntimeError):
        pyoracle.decode(bytes(enc[: len(enc) // 2]), len(data) + 64)  # truncated: UnexpectedEof
    with pytest.raises(RuntimeError):
        pyoracle.decode(bytes([0xFF, 0xFF, 0xFF, 0xFF, 0x7F]) + bytes(100), 1000)  # chunk length >= 3*PREMATCH: InvalidData


def test_chunk_level_interface_equals_stream_interface():
    """The operator-level restatement (lz.rs:75-95) driven like lib.rs:58-92 reproduces the stream-level output."""
    data = datagen.text_like(150_000, seed=71) + datagen.random_bytes(20_000, seed=72)
    cfg = pyoracle.level_cfg(1)
    enc = pyoracle.ChunkEncoder()
    window = np.zeros(pyoracle.BLOCK_SIZE + 2 * pyoracle.SENTINEL, dtype=np.uint8)
    src = np.frombuffer(data, dtype=np.uint8)
    window[pyoracle.SENTINEL + pyoracle.PREMATCH: pyoracle.SENTINEL + pyoracle.PREMATCH + src.size] = src
    out, spos = bytearray(), pyoracle.PREMATCH
    while spos < pyoracle.PREMATCH + src.size:
        spos, payload = enc.encode(cfg, window, pyoracle.PREMATCH + src.size, spos)
        n, b = len(payload), bytearray()
        while n >= 128:
            b.append(128 + n % 128)
            n //= 128
        b.append(n)
        out += bytes(b) + payload
        pre, post = enc.last_items()
        assert len(pre) == len(post) > 0 and (pre["ctx"] == post["ctx"]).all() and (post["symbol"] <= 388).all()
    out += b"\x00"
    assert bytes(out) == pyoracle.encode(data, 1)


def test_round_trip_property_random_structures():
    """Seeded pseudo-random structured inputs (mixtures of runs, copies and noise) round-trip at every level."""
    rng = np.random.default_rng(2024)
    for trial in range(60):
        parts = []
        for _ in range(int(rng.integers(1, 12))):
            kind = int(rng.integers(0, 4))
            if kind == 0:
                parts.append(bytes(rng.integers(0, 256, size=int(rng.integers(1, 400)), dtype=np.uint8)))
            elif kind == 1:
                parts.append(bytes([int(rng.integers(0, 256))]) * int(rng.integers(1, 700)))
            elif kind == 2 and parts:
                src = b"".join(parts)
                o = int(rng.integers(0, len(src)))
                parts.append(src[o:o + int(rng.integers(1, 300))])
            else:
                parts.append(bytes(rng.integers(97, 123, size=int(rng.integers(1, 60)), dtype=np.uint8)) * int(rng.integers(1, 9)))
        data = b"".join(parts)
        level = trial % 3
        assert pyoracle.decode(pyoracle.encode(data, level), len(data) + 64) == data


# ---------------------------------------------------------------------------------------------------------------------
# Second restatement (oracle/pyref.py: plain Python written from src/*.rs, not from the C oracle) vs the C oracle.
# Two independent readings of the reference that disagree anywhere in parse, rank coding, Huffman build or framing cannot
# produce the same bytes; agreement is not a pin against real `orz` output (tests/golden/ref/ is), but it removes
# "one misreading, faithfully tested against itself" as a failure mode.
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("gen,n,level", [("text_like", 60_000, 2), ("text_like", 40_000, 0), ("low_entropy", 50_000, 1),
                                         ("random_bytes", 12_000, 0), ("long_chains", 40_000, 2)])
def test_second_restatement_agrees(gen, n, level):
    from oracle import pyref
    data = getattr(datagen, gen)(n, seed=77)
    assert pyref.encode(data, level) == pyoracle.encode(data, level)


def test_second_restatement_agrees_tiny_and_custom_cfg():
    from oracle import pyref
    for data in (b"", b"a", b"ab", b"abcd", b"aaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaa", b"the quick brown fox jumps over the lazy dog. " * 30):
        for level in (0, 1, 2):
            assert pyref.encode(data, level) == pyoracle.encode(data, level)
    data = datagen.text_like(30_000, seed=5)
    for cfg in ((70, 40, 33), (1, 1, 1), (0, 0, 0)):
        assert pyref.encode(data, cfg) == pyoracle.encode(data, cfg=pyoracle.Cfg(*cfg))


def test_second_restatement_agrees_enwik8_prefix():
    from oracle import pyref
    e8 = datagen.enwik8(1_200_000)
    if e8 is None:
        pytest.skip("data/enwik8.xz not present")
    for level in (0, 2):  # > 2^18 items: rank tables, three Huffman tables, 240-byte matches, WORD symbols
        assert pyref.encode(e8, level) == pyoracle.encode(e8, level)


def test_second_restatement_agrees_across_block_slides(tmp_path):
    """Scale model (128 KiB blocks, 3000-item chunks, as tests/simt): block slides, `forward`, stale tail bytes, chunk cuts —
    pyref with those two constants vs the C oracle compiled with the same two constants."""
    import subprocess
    from oracle import pyref
    simt = os.path.join(ROOT, "tests", "simt")
    subprocess.check_call(["make", "-C", simt, "orz_oracle_small"], stdout=subprocess.DEVNULL)
    data = datagen.text_like(200_000, seed=61) + datagen.low_entropy(100_000, seed=5)
    src, out = tmp_path / "in", tmp_path / "want"
    src.write_bytes(data)
    for level in (0, 2):
        subprocess.check_call([os.path.join(simt, "orz_oracle_small"), "encode", str(level), str(src), str(out)], stderr=subprocess.DEVNULL)
        assert pyref.encode(data, level, block_bits=17, chunk_items=3000) == out.read_bytes()


# ---------------------------------------------------------------------------------------------------------------------
# Hand-derived PARSE vectors (derived on paper from lz.rs:131-235 / matcher.rs:135-228 before the oracle was asked; the
# derivations are in the docstrings).  SURVEY App. D's two vectors only reach the empty and single-literal paths; these
# reach a match with its length code, WORD symbols, and a lazy-match rejection.
# ---------------------------------------------------------------------------------------------------------------------
def _pre_rank_items(data, cfg):
    enc = pyoracle.ChunkEncoder()
    window = np.zeros(pyoracle.BLOCK_SIZE + 2 * pyoracle.SENTINEL, dtype=np.uint8)
    src = np.frombuffer(data, dtype=np.uint8)
    window[pyoracle.SENTINEL + pyoracle.PREMATCH: pyoracle.SENTINEL + pyoracle.PREMATCH + src.size] = src
    spos, _ = enc.encode(cfg, window, pyoracle.PREMATCH + src.size, pyoracle.PREMATCH)
    assert spos == pyoracle.PREMATCH + src.size
    return enc.last_items()[0]


def test_hand_derived_parse_match():
    """b"0123456789abcdef" * 2 + b"XYZW", -l2.  Positions 0..15: no candidates, literals.  16 ('0'): bucket hash1(15) = 'f'|alnum('e')<<7
    holds nothing with this dword -> literal; 17 ('1'): bucket '0'|alnum('f')<<7 = 0xB0, the first '1' sits in bucket '0'|0 = 0x30 ->
    literal.  18 ('2'): bucket '1'|alnum('0')<<7 = 0xB1 holds the node of position 2 (same bucket, same dword "2345"): LCP =
    "23456789abcdef" = 14 (then '0' vs 'X'), 18+14 < 36, expected of that node 0 -> no early exit.  reduced offset 0 -> roid 0,
    robitlen 0 (< 8) -> lazy_len1 = 16; position 19 matches 13 bytes, position 20 12 bytes -> no lazy match.  encoded_match_len:
    14 > max(expected 0, 4) -> 14 - max(lenmin 0, 4) = 10, lenid 5, symbol 256 + 0*6 + 5 = 261; ctx = 0xB1 | after_literal << 8 = 433.
    32 ('X'): words[hash2(31)] was written by the literal at 17 with ('0','1') -> unlikely '0', not matched -> literal; Y, Z, W literals."""
    pre = _pre_rank_items(b"0123456789abcdef" * 2 + b"XYZW", pyoracle.level_cfg(2))
    assert list(pre["symbol"]) == [ord(c) for c in "0123456789abcdef01"] + [261] + [ord(c) for c in "XYZW"]
    m = pre[18]
    assert (m["ctx"], m["robitlen"], m["robits"], m["eml"], m["after_literal"], m["is_match"]) == (433, 0, 0, 10, 1, 1)
    assert pre[19]["unlikely"] == ord("0") and pre[19]["after_literal"] == 0


def test_hand_derived_parse_word_symbols():
    """b"0123456789abcdef" * 2 with LZCfg(0,0,0): find_match never enters its loop (matcher.rs:149), so every step is a literal or a WORD.
    The literal at q stores words[hash2(q-2)] = (b[q-1], b[q]) (lz.rs:233); the lookup at p reads words[hash2(p-1)] (lz.rs:132).  In the
    second copy the keys line up from p = 19 on (key ('2','1',alnum '0') was stored by the literal at 4 with ('3','4')): WORD at
    19, 21, 23, 25, 27, 29 (lz.rs:218-226: no words update, after_literal = false).  16, 17, 18 miss because the third key component
    (alnum of the byte three back) differs from the first copy's, where that byte lay before the stream.  31 is the last byte: spos + 1 <
    len fails -> literal 'f'."""
    pre = _pre_rank_items(b"0123456789abcdef" * 2, pyoracle.Cfg(0, 0, 0))
    assert list(pre["symbol"]) == [ord(c) for c in "0123456789abcdef012"] + [388] * 6 + [ord("f")]
    assert [int(u) for u in pre["unlikely"][19:25]] == [ord(c) for c in "3579bd"]
    assert [int(a) for a in pre["after_literal"][19:26]] == [1, 0, 0, 0, 0, 0, 0]


def test_hand_derived_parse_lazy_rejection():
    """b".9ABCD!~-zABCDEFGH?~,9ABCDEFGH#qwertyuiop", -l2.  First occurrence ".9ABCD!": literals.  Second "-zABCDEFGH": 'A','B' literals
    (their dwords differ from the first occurrence's: "BCD!" vs "BCDE"); at 'C' (12) words[hash2(11)] — key ('B','A',alnum 'z') = the key
    the literal at 5 stored ('C','D') under (('B','A',alnum '9')) — predicts "CD": WORD; E F G H ? ~ , 9 literals (",9": the '9' sits in
    bucket ','|0, the first one in '.'|0: no match).  22 ('A'): bucket '9'|0 = 0x39 holds position 2, "ABCD" then '!' vs 'E': match 4,
    reduced offset 0 -> lazy_len1 = 4+1+1 = 6.  Lazy probe 1: position 23 in bucket 'A'|alnum('9')<<7 = 0xC1, entry of "BCDE" = position 11
    (second occurrence, same bucket because 'z' is alphanumeric too), "BCDEFG" equal -> has_lazy_match -> lazy_match_id 1 -> literal 'A'
    (lz.rs:218 excludes WORD).  23 ('B'): match with position 11, LCP "BCDEFGH" = 7; lazy_len1 = 9: position 24 vs 12 shares 6, and with
    last_word_matched (the literal at 4 stored ('B','C') under the key of 23) lazy_len2 = 8: position 25's bucket 'C'|1<<7 has no
    "DEFG" node (13 was inside the WORD) -> match emitted: 7 > 4 -> encoded_match_len 3, symbol 259, ctx 0xC1|1<<8 = 449, unlikely 'B'."""
    pre = _pre_rank_items(b".9ABCD!~-zABCDEFGH?~,9ABCDEFGH#qwertyuiop", pyoracle.level_cfg(2))
    want = [ord(c) for c in ".9ABCD!~-zAB"] + [388] + [ord(c) for c in "EFGH?~,9A"] + [259] + [ord(c) for c in "#qwertyuiop"]
    assert list(pre["symbol"]) == want
    m = pre[22]
    assert (m["ctx"], m["unlikely"], m["eml"], m["robitlen"], m["after_literal"], m["is_match"]) == (449, ord("B"), 3, 0, 1, 1)


def test_reference_golden_vectors_when_present():
    """tests/golden/ref/*.json are written by tools/make_ref_golden.sh on a machine with nightly Rust: sha256/len of REAL `orz encode`
    output.  None exist in this tree (no Rust toolchain in the authoring container nor on the GPU box): until that script has run
    somewhere, parity with real orz bytes stays unpinned and this test skips."""
    ref_dir = os.path.join(ROOT, "tests", "golden", "ref")
    files = [f for f in sorted(os.listdir(ref_dir)) if f.endswith(".json")] if os.path.isdir(ref_dir) else []
    if not files:
        pytest.skip("no reference-produced vectors (tools/make_ref_golden.sh needs nightly Rust)")
    from tools import ref_golden_cases
    for f in files:
        rec = json.load(open(os.path.join(ref_dir, f)))
        data = ref_golden_cases.case_bytes(rec["case"])
        assert hashlib.sha256(data).hexdigest() == rec["sha256_in"]
        enc = pyoracle.encode(data, int(rec["level"]))
        assert (len(enc), hashlib.sha256(enc).hexdigest()) == (rec["len"], rec["sha256"]), f


def test_parse_diff_tool_locates_the_first_differing_step():
    """tools/diff_parse.py (the oracle's chunk decoder with its item trace on) is how a stage A timing hole was localised in round 2:
    two streams of the same input are decoded side by side and the first parse step that differs is reported with its input
    position.  Here: the -l1 and -l2 streams of one text (same literals / matches until the deeper search finds a longer match)."""
    from tools import diff_parse
    data = datagen.text_like(200_000, seed=5)
    a, b = pyoracle.encode(data, 2), pyoracle.encode(data, 1)
    assert diff_parse.first_difference(a, a) is None
    d = diff_parse.first_difference(a, b)
    assert d is not None
    ci, k, before, ra, rb = d
    assert ci == 0 and before == 0 and 0 < k < min(len(ra), len(rb))
    assert (ra[:k] == rb[:k]).all() and (ra[k] != rb[k]).any()
    assert int(ra[k][0]) == int(rb[k][0])  # both parses stand at the same input position when they part
    # every traced position is where the previous item ended: literal +1, WORD +2, match +length
    adv = np.where(ra[:, 1] < 256, 1, np.where(ra[:, 2] > 0, ra[:, 2], 2))
    assert (ra[1:, 0] == ra[:-1, 0] + adv[:-1]).all() and int(ra[-1, 0] + adv[-1]) == len(data)
