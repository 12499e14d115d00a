"""orz_b200 — B200-native (sm_100a CUDA) ROLZ codec, a drop-in for richox/orz's encode/decode hot path.

Public surface mirrors the reference crate (src/lib.rs:22-24,58,94): encode, decode, LZCfg, plus the
operator interface LZEncoder / LZDecoder and the batch extension.  No CPU fallback exists.
"""
from ._lib import LZCfg, OrzError, LIB_PATH, SYMBOLS  # noqa: F401
from .codec import (Batch, LZDecoder, LZEncoder, decode, decode_batch, encode_batch_multi, decode_segmented, decode_with_operators,  # noqa: F401
                    device_count, encode, encode_batch, encode_bound, encode_segmented, encode_with_operators, segment_streams)
