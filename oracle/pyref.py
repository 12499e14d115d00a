"""Second, independent restatement of richox/orz's encoder (TEST INFRASTRUCTURE, like orz_oracle.c).

Written directly from the Rust sources (richox/orz @29414c6: src/lib.rs, lz.rs, matcher.rs, symrank.rs, huffman.rs, coder.rs,
mem.rs, ioutil.rs) WITHOUT reference to oracle/orz_oracle.c, in plain Python (lists, bytearrays, heapq), so that a
misreading of the reference in one restatement shows up as a difference between the two.  tests/test_oracle.py diffs the two
on tiny inputs, text, low-entropy data, the first megabytes of enwik8 and a multi-block scale model.  It pins nothing against real
`orz` bytes (no Rust toolchain here — see tools/make_ref_golden.sh for that); it is slow (tens of KB/s) and encode-only.

Every function cites the reference lines it follows.  Names follow the reference.
"""
import heapq

LZ_CHUNK_SIZE = 1 << 20          # lib.rs:32
LZ_MATCH_MAX_LEN = 240           # lib.rs:33
LZ_MATCH_MIN_LEN = 4             # lib.rs:34
LZ_MF_BUCKET_ITEM_SIZE = 4094    # lz.rs:24
LZ_MF_BUCKET_ITEM_HASH_SIZE = int(LZ_MF_BUCKET_ITEM_SIZE * 1.13) | 1  # matcher.rs:18  (4627)
LZ_ROID_SIZE = 22                # lz.rs:27
LZ_LENID_SIZE = 6                # lz.rs:28
SYMRANK_NUM_SYMBOLS = 256 + LZ_ROID_SIZE * LZ_LENID_SIZE + 1  # lz.rs:25
WORD_SYMBOL = SYMRANK_NUM_SYMBOLS - 1                         # lz.rs:29
SENTINEL = LZ_MATCH_MAX_LEN * 2  # lib.rs:54
M32 = 0xFFFFFFFF


def _roid_encoding_array():  # lz.rs:494-514
    encs = [(0, 0, 0)] * LZ_MF_BUCKET_ITEM_SIZE
    base = current_id = enc_idx = 0
    while base < LZ_MF_BUCKET_ITEM_SIZE:
        bit_len = current_id // 2  # lz.rs:532-534
        rest_bits = 0
        while rest_bits != (1 << bit_len):
            if base < LZ_MF_BUCKET_ITEM_SIZE:
                encs[enc_idx] = (current_id, bit_len, rest_bits)
                enc_idx += 1
                base += 1
            rest_bits += 1
        current_id += 1
    return encs


ROID = _roid_encoding_array()


# ---------------------------------------------------------------------------------------------- coder.rs
class Encoder:
    """coder.rs:12-89 with BitBuffer (coder.rs:159-217) folded in: value/len, big-endian u32 flushes."""

    def __init__(self):
        self.out = bytearray()
        self.value = 0
        self.len = 0

    def _reserve(self):  # coder.rs:84-88, 198-206
        if self.len >= 32:
            self.len -= 32
            w = (self.value >> self.len) & M32
            self.out += w.to_bytes(4, "big")  # swap_bytes of the little-endian store
            self.value &= (1 << self.len) - 1

    def _put(self, n, v):  # coder.rs:183-187
        self.value = ((self.value << n) ^ v) & 0xFFFFFFFFFFFFFFFF
        self.len += n

    def varint(self, v):  # coder.rs:27-38
        while True:
            self._reserve()
            has_next = v > 1
            self._put(2, (v & 1) | (int(has_next) << 1))
            v >>= 1
            if not has_next:
                break

    def raw_bits(self, bits, n):  # coder.rs:40-43
        self._reserve()
        self._put(n, bits)

    def huffman_table(self, code_lens):  # coder.rs:45-67
        max_code_len = max(code_lens)
        self.varint(max_code_len)
        last_sym = None
        for sym, cl in enumerate(code_lens):
            if cl > 0:
                self.varint(sym + 1 if last_sym is None else sym - last_sym)
                self.varint(max_code_len - cl)
                last_sym = sym
        self.varint(0)

    def huffman_sym(self, enc, sym):  # coder.rs:69-73
        self._reserve()
        bits, cl = enc[sym]
        self._put(cl, bits)

    def finish(self):  # coder.rs:75-82, 208-216
        self._reserve()
        if self.len > 0:
            self._put(32 - self.len, 0)
            while self.len > 0:
                self.out.append((self.value >> (self.len - 8)) & 0xFF)
                self.len -= min(self.len, 8)
        return bytes(self.out)


# ---------------------------------------------------------------------------------------------- huffman.rs
def huffman_code_lens(sym_weights, max_code_len=15):
    """huffman.rs:27-111.  The BinaryHeap holds HeapNode{weight,index} whose PartialOrd is reversed (huffman.rs:34-38) and std's
    heap sifts with <= / >=, i.e. it pops the SMALLEST (weight, index): heapq on tuples."""
    n = len(sym_weights)
    weights = list(sym_weights)
    while True:
        heap = [(weights[i], i) for i in range(n) if sym_weights[i] > 0]
        heapq.heapify(heap)
        if len(heap) <= 1:
            lens = [0] * n
            if heap:
                lens[heap[0][1]] = 1
            return lens
        child = {}
        nxt = n
        while len(heap) > 1:
            w1, i1 = heapq.heappop(heap)
            w2, i2 = heapq.heappop(heap)
            child[nxt] = (i1, i2)
            heapq.heappush(heap, (w1 + w2, nxt))
            nxt += 1
        lens = [0] * nxt
        for i in range(nxt - 1, n - 1, -1):
            c1, c2 = child[i]
            lens[c1] = lens[i] + 1
            lens[c2] = lens[i] + 1
        lens = lens[:n]
        cur = max(lens)
        if cur > max_code_len:
            shrink = 1 << (cur - max_code_len)
            weights = [max(w // shrink, 1) if w > 0 else w for w in weights]
            continue
        return lens


def huffman_encoding(code_lens):  # huffman.rs:119-141
    enc = [(0, 0)] * len(code_lens)
    bits = 0
    cur = 1
    for sym in sorted((i for i in range(len(code_lens)) if code_lens[i] > 0), key=lambda s: (code_lens[s], s)):
        shift = code_lens[sym] - cur
        if shift > 0:
            bits <<= shift
            cur += shift
        enc[sym] = (bits, code_lens[sym])
        bits += 1
    return enc


# ---------------------------------------------------------------------------------------------- symrank.rs
class SymRankCoder:
    def __init__(self):  # symrank.rs:22-29
        self.value_array = [0] * SYMRANK_NUM_SYMBOLS
        self.index_array = [0] * SYMRANK_NUM_SYMBOLS
        self.encoded_cnt = 0
        self.encoded_idx_sum = 1000000

    def init(self, values):  # symrank.rs:31-36
        for i, v in enumerate(values):
            self.value_array[i] = v
            self.index_array[v] = i

    def clone(self):
        c = SymRankCoder()
        c.value_array = list(self.value_array)
        c.index_array = list(self.index_array)
        c.encoded_cnt = self.encoded_cnt
        c.encoded_idx_sum = self.encoded_idx_sum
        return c

    def encode(self, v, vunlikely):  # symrank.rs:38-47
        i = self.index_array[v]
        iunlikely = self.index_array[vunlikely]
        self._update(v, i)
        if i == iunlikely:
            return SYMRANK_NUM_SYMBOLS - 1
        return i - (1 if i > iunlikely else 0)

    def _update(self, v, i):  # symrank.rs:61-97 (u32 / u16 release arithmetic)
        if self.encoded_cnt > SYMRANK_NUM_SYMBOLS:
            self.encoded_cnt = (self.encoded_cnt * 9 & M32) // 10
            self.encoded_idx_sum = (self.encoded_idx_sum * 9 & M32) // 10
        self.encoded_cnt = (self.encoded_cnt + 1) & M32
        self.encoded_idx_sum = (self.encoded_idx_sum + i) & M32
        sub = (i // 16 + ((self.encoded_idx_sum // 16 // self.encoded_cnt) & 0xFFFF)) & 0xFFFF  # `as u16`, wrapping u16 add
        next_i = max(i - sub if i > sub else 0, i // 2)                                          # saturating_sub(..).max(i / 2)
        n = i - next_i
        va, ia = self.value_array, self.index_array
        if n == 0:
            return
        if n == 1:
            ni1 = next_i
            nv1 = va[ni1]
            ia[v] = ni1
            va[i] = nv1
            ia[nv1] = i
            va[ni1] = v
        else:
            ni2 = next_i
            ni1 = next_i + n // 2
            nv1 = va[ni1]
            nv2 = va[ni2]
            va[i] = nv1
            ia[nv1] = i
            va[ni1] = nv2
            ia[nv2] = ni1
            va[ni2] = v
            ia[v] = ni2


# ---------------------------------------------------------------------------------------------- matcher.rs
def hash_dword(buf, pos):  # matcher.rs:255-263
    return ((buf[pos] * 131313131 ^ 797) + (buf[pos + 1] * 1313131 ^ 79797) + (buf[pos + 2] * 13131 ^ 7979797) +
            (buf[pos + 3] * 131 ^ 797979797)) & M32


def common_prefix(buf, p1, p2, max_len):  # mem.rs:41-51: 16-byte steps, first differing byte
    for l in range(0, max_len, 16):
        a = buf[p1 + l:p1 + l + 16]
        b = buf[p2 + l:p2 + l + 16]
        if a != b:
            for k in range(16):
                if a[k] != b[k]:
                    return l + k
    return max_len


class Bucket:  # matcher.rs:28-100
    def __init__(self):
        self.pos = [0] * LZ_MF_BUCKET_ITEM_SIZE
        self.len_min = [0] * LZ_MF_BUCKET_ITEM_SIZE
        self.len_expected = [0] * LZ_MF_BUCKET_ITEM_SIZE
        self.head = 0

    def update(self, pos, reduced_offset, match_len):  # matcher.rs:62-80
        new_head = (self.head + 1) % LZ_MF_BUCKET_ITEM_SIZE
        if match_len >= LZ_MATCH_MIN_LEN:
            ni = (self.head + LZ_MF_BUCKET_ITEM_SIZE - reduced_offset) % LZ_MF_BUCKET_ITEM_SIZE
            if self.len_min[ni] <= match_len:
                self.len_min[ni] = min(match_len + 1, 127)
        self.pos[new_head] = pos
        self.len_min[new_head] = 0
        self.len_expected[new_head] = match_len
        self.head = new_head

    def forward(self, forward_len):  # matcher.rs:82-87 (saturating_sub)
        self.pos = [p - forward_len if p > forward_len else 0 for p in self.pos]


class BucketMatcher:  # matcher.rs:102-229
    def __init__(self):
        self.heads = [-1] * LZ_MF_BUCKET_ITEM_HASH_SIZE
        self.nexts = [-1] * LZ_MF_BUCKET_ITEM_SIZE

    def update(self, bucket, buf, pos):  # matcher.rs:115-121
        entry = hash_dword(buf, pos) % LZ_MF_BUCKET_ITEM_HASH_SIZE
        self.nexts[bucket.head] = self.heads[entry]
        self.heads[entry] = bucket.head

    def forward(self, bucket):  # matcher.rs:123-133
        self.heads = [-1 if (h != -1 and bucket.pos[h] == 0) else h for h in self.heads]
        self.nexts = [-1 if (x != -1 and bucket.pos[x] == 0) else x for x in self.nexts]

    def find_match(self, bucket, buf, buf_len, pos, match_depth):  # matcher.rs:135-192
        entry = hash_dword(buf, pos) % LZ_MF_BUCKET_ITEM_HASH_SIZE
        node_index = self.heads[entry]
        if node_index == -1:
            return None
        max_len = LZ_MATCH_MIN_LEN - 1
        max_match_len_min = LZ_MATCH_MIN_LEN
        max_match_len_expected = LZ_MATCH_MIN_LEN
        max_node_index = 0
        node_pos = bucket.pos[node_index]
        max_len_dword = buf[pos + max_len - 3:pos + max_len + 1]
        for _ in range(match_depth):
            if buf[node_pos + max_len - 3:node_pos + max_len + 1] == max_len_dword:
                lcp = common_prefix(buf, node_pos, pos, LZ_MATCH_MAX_LEN)
                if lcp > max_len:
                    max_match_len_min = bucket.len_min[node_index]
                    max_match_len_expected = bucket.len_expected[node_index]
                    max_len = lcp
                    max_node_index = node_index
                    max_len_dword = buf[pos + max_len - 3:pos + max_len + 1]
                if lcp == LZ_MATCH_MAX_LEN:
                    break
                if max_match_len_expected > 0 and lcp > max_match_len_expected:
                    break
            node_index = self.nexts[node_index]
            if node_index == -1:
                break
            node_pos_next = bucket.pos[node_index]
            if node_pos <= node_pos_next:
                break
            node_pos = node_pos_next
        if max_len >= LZ_MATCH_MIN_LEN and pos + max_len < buf_len:
            return ((bucket.head + LZ_MF_BUCKET_ITEM_SIZE - max_node_index) % LZ_MF_BUCKET_ITEM_SIZE, max_len,
                    max(max_match_len_expected, LZ_MATCH_MIN_LEN), max(max_match_len_min, LZ_MATCH_MIN_LEN))
        return None

    def has_lazy_match(self, bucket, buf, pos, min_match_len, depth):  # matcher.rs:194-228, mem.rs:55-70
        want = buf[pos:pos + min_match_len]
        entry = hash_dword(buf, pos) % LZ_MF_BUCKET_ITEM_HASH_SIZE
        node_index = self.heads[entry]
        if node_index == -1:
            return False
        node_pos = bucket.pos[node_index]
        for _ in range(depth):
            if buf[node_pos:node_pos + min_match_len] == want:
                return True
            node_index = self.nexts[node_index]
            if node_index == -1:
                break
            node_pos_next = bucket.pos[node_index]
            if node_pos <= node_pos_next:
                break
            node_pos = node_pos_next
        return False


# ---------------------------------------------------------------------------------------------- lz.rs
def _is_alnum(c):
    return (48 <= c <= 57) or (65 <= c <= 90) or (97 <= c <= 122)


def hash1(buf, pos):  # lz.rs:481-486
    return (buf[pos] & 0x7F) | (int(_is_alnum(buf[pos - 1])) << 7)


def hash2(buf, pos):  # lz.rs:488-492
    return (buf[pos] & 0x7F) | (hash1(buf, pos - 1) << 7)


class LZEncoder:
    def __init__(self):  # lz.rs:58-80
        self.buckets = [Bucket() for _ in range(256)]
        self.matchers = [BucketMatcher() for _ in range(256)]
        self.symranks = [SymRankCoder() for _ in range(512)]
        self.words = [(0, 0)] * 32768
        self.first_block = True
        self.after_literal = True

    def forward(self, forward_len):  # lz.rs:82-87
        for b, m in zip(self.buckets, self.matchers):
            b.forward(forward_len)
            m.forward(b)

    def encode(self, cfg, buf, base, sbuf_len, spos, chunk_items=LZ_CHUNK_SIZE):
        """lz.rs:89-345.  `buf` is the whole sbvec_buf, window position p lives at buf[base + p] (base = the leading sentinel),
        so the reference's reads before position 0 and past sbuf.len() see the same bytes.  Returns (spos, payload)."""
        depth0, depth1, depth2 = cfg
        items = []  # (symbol, ctx, unlikely, after_literal, robitlen, robits, encoded_match_len) ; robitlen None for Symbol items
        B = base
        while spos < sbuf_len and len(items) < chunk_items:  # lz.rs:131
            last_word_expected = self.words[hash2(buf, B + spos - 1)]
            last_word_matched = (buf[B + spos], buf[B + spos + 1]) == last_word_expected
            c = hash1(buf, B + spos - 1)
            symrank_context = c | (int(self.after_literal) << 8)
            symrank_unlikely = last_word_expected[0]
            lazy_match_id = 0
            m = self.matchers[c].find_match(self.buckets[c], buf, B + sbuf_len, B + spos, depth0)
            if m is not None:
                reduced_offset, match_len, match_len_expected, match_len_min = m
                roid, robitlen, robits = ROID[reduced_offset]
                if match_len < LZ_MATCH_MAX_LEN // 2:  # lz.rs:151-171
                    lazy_len1 = match_len + 1 + int(robitlen < 8)
                    lazy_len2 = lazy_len1 - int(last_word_matched)
                    c1 = hash1(buf, B + spos)
                    if self.matchers[c1].has_lazy_match(self.buckets[c1], buf, B + spos + 1, lazy_len1, depth1):
                        lazy_match_id = 1
                    else:
                        c2 = hash1(buf, B + spos + 1)
                        if self.matchers[c2].has_lazy_match(self.buckets[c2], buf, B + spos + 2, lazy_len2, depth2):
                            lazy_match_id = 2
                if lazy_match_id == 0:  # lz.rs:173-209
                    if match_len > match_len_expected:
                        eml = match_len - match_len_min
                    elif match_len < match_len_expected:
                        eml = match_len - match_len_min + 1
                    else:
                        eml = 0
                    eml &= 0xFF  # `as u8` of a wrapping usize
                    lenid = min(LZ_LENID_SIZE - 1, eml)
                    items.append((256 + roid * LZ_LENID_SIZE + lenid, symrank_context, symrank_unlikely, self.after_literal, robitlen,
                                  robits, eml))
                    # positions are stored window-relative (matcher.rs:76): the bucket sees spos, the byte reads use B + pos
                    self.buckets[c].update(B + spos, reduced_offset, match_len)
                    self.matchers[c].update(self.buckets[c], buf, B + spos)
                    spos += match_len
                    self.after_literal = False
                    self.words[hash2(buf, B + spos - 3)] = (buf[B + spos - 2], buf[B + spos - 1])
                    continue
            self.buckets[c].update(B + spos, 0, 0)  # lz.rs:210-215
            self.matchers[c].update(self.buckets[c], buf, B + spos)
            if spos + 1 < sbuf_len and lazy_match_id != 1 and last_word_matched:  # lz.rs:218-226
                items.append((WORD_SYMBOL, symrank_context, symrank_unlikely, self.after_literal, None, 0, 0))
                spos += 2
                self.after_literal = False
            else:  # lz.rs:227-235
                items.append((buf[B + spos], symrank_context, symrank_unlikely, self.after_literal, None, 0, 0))
                spos += 1
                self.after_literal = True
                self.words[hash2(buf, B + spos - 3)] = (buf[B + spos - 2], buf[B + spos - 1])

        enc = Encoder()
        if self.first_block:  # lz.rs:238-265
            counts = [0] * SYMRANK_NUM_SYMBOLS
            for it in items:
                counts[it[0]] += 1
            num_counted = sum(1 for x in counts if x > 1)
            vs = sorted(range(SYMRANK_NUM_SYMBOLS), key=lambda i: -max(counts[i], 1))  # stable, Reverse(count.max(1))
            enc.varint(num_counted)
            for s in vs[:num_counted]:
                enc.raw_bits(s, 9)
            initial = SymRankCoder()
            initial.init(vs)
            self.symranks = [initial.clone() for _ in range(512)]
            self.first_block = False
        enc.varint(min(spos, sbuf_len))  # lz.rs:268-269
        enc.varint(len(items))
        w1 = [[0] * SYMRANK_NUM_SYMBOLS, [0] * SYMRANK_NUM_SYMBOLS]  # lz.rs:272-305
        w2 = [0] * LZ_MATCH_MAX_LEN
        ranked = []
        for sym, ctx, unl, al, robitlen, robits, eml in items:
            r = self.symranks[ctx].encode(sym, unl)
            w1[int(al)][r] += 1
            if robitlen is not None and eml >= LZ_LENID_SIZE - 1:
                w2[eml] += 1
            ranked.append(r)
        t1 = [huffman_code_lens(w1[0]), huffman_code_lens(w1[1])]  # lz.rs:306-318
        t2 = huffman_code_lens(w2)
        enc.huffman_table(t1[0])
        enc.huffman_table(t1[1])
        enc.huffman_table(t2)
        h1 = [huffman_encoding(t1[0]), huffman_encoding(t1[1])]
        h2 = huffman_encoding(t2)
        for (sym, ctx, unl, al, robitlen, robits, eml), r in zip(items, ranked):  # lz.rs:320-342
            enc.huffman_sym(h1[int(al)], r)
            if robitlen is not None:
                enc.raw_bits(robits, robitlen)
                if eml >= LZ_LENID_SIZE - 1:
                    enc.huffman_sym(h2, eml)
        return spos, enc.finish()


def write_len(n):  # ioutil.rs:79-88
    out = bytearray()
    while n >= 128:
        out.append(128 + n % 128)
        n //= 128
    out.append(n)
    return bytes(out)


LEVELS = {0: (5, 3, 2), 1: (15, 9, 6), 2: (45, 27, 18)}  # main.rs:97-102


def encode(data: bytes, level=2, block_bits=25, chunk_items=LZ_CHUNK_SIZE) -> bytes:
    """lib.rs:58-92.  block_bits / chunk_items other than 25 / 2^20 give the scale model of tests/simt (the reference's constants
    LZ_BLOCK_SIZE = 2^25 - 1 and LZ_CHUNK_SIZE = 2^20 shrunk so that block slides and chunk cuts happen within kilobytes)."""
    cfg = LEVELS[level] if isinstance(level, int) else tuple(level)
    block_size = (1 << block_bits) - 1
    prematch = block_size // 2
    lzenc = LZEncoder()
    # bucket positions are stored as buf indices (window position + SENTINEL); `forward` subtracts window distances, and a
    # window position 0 must read as "dead": keep the reference's window-relative values by storing B + pos with B subtracted
    # in forward().  Simpler and exact: run the matcher on window-relative positions by giving it a view that starts at SENTINEL.
    buf = bytearray(block_size + 2 * SENTINEL)
    out = bytearray()
    view = _WindowView(buf, SENTINEL)
    pos_in = 0
    while True:
        chunk = data[pos_in:pos_in + (block_size - prematch)]  # lib.rs:72 read_repeatedly into sbvec[PREMATCH..]
        if not chunk:
            break
        pos_in += len(chunk)
        buf[SENTINEL + prematch:SENTINEL + prematch + len(chunk)] = chunk
        spos = prematch
        sbuf_len = prematch + len(chunk)
        while spos < sbuf_len:  # lib.rs:76-82
            spos, payload = lzenc.encode(cfg, view, 0, sbuf_len, spos, chunk_items)
            out += write_len(len(payload))
            out += payload
        buf[SENTINEL:SENTINEL + prematch] = buf[SENTINEL + block_size - prematch:SENTINEL + block_size]  # lib.rs:83 copy_within
        lzenc.forward(block_size - prematch)  # lib.rs:84
    out += write_len(0)  # lib.rs:89
    return bytes(out)


class _WindowView:
    """sbvec = &mut sbvec_buf[SENTINEL..][..LZ_BLOCK_SIZE] (lib.rs:69): indexable by window position, negative positions and positions
    past the slice land in the sentinels exactly as the reference's unchecked pointer reads do."""

    def __init__(self, buf, base):
        self.buf = buf
        self.base = base

    def __getitem__(self, i):
        if isinstance(i, slice):
            return self.buf[self.base + i.start:self.base + i.stop]
        return self.buf[self.base + i]
