"""Subprocess body of tests/test_gpu_stage_a_variants.py: one stage A variant (switches given as KEY=VALUE arguments)
against the oracle on a two-chunk input and on a two-block input.  A separate process, so that a fault in an
experimental kernel cannot poison the CUDA context of the test session."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

for a in sys.argv[1:]:
    k, v = a.split("=", 1)
    os.environ[k] = v

import orz_b200  # noqa: E402
from oracle import pyoracle  # noqa: E402
from tests import datagen  # noqa: E402

e8 = datagen.enwik8(6_000_000)
two_chunks = e8 if e8 is not None else datagen.text_like(6_000_000, seed=91)   # > 2^20 items: a chunk cut (lz.rs:128-131)
two_blocks = (datagen.text_like(3_000_000, seed=92) + datagen.low_entropy(2_000_000, seed=93)) * 4  # 20 MB: slide + forward
deep = datagen.long_chains(1_500_000, seed=94)  # chain walks to the depth limit, past one ring-record snapshot
for name, data, level in (("two chunks", two_chunks, 2), ("two blocks", two_blocks, 1), ("long chains", deep, 2)):
    got = orz_b200.encode(data, level)
    want = pyoracle.encode(data, level)
    if got != want:
        print(f"MISMATCH {name}: {len(got)} vs {len(want)} bytes")
        sys.exit(1)
print("identical")
