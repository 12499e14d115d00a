"""Copies the reference's test corpus (test/enwik8.xz) into data/ so that it travels to the GPU box.

data/enwik8.xz is git-ignored (26 MB of input data, not source).  Without it bench.py and the tests fall back to the
committed order-2 Markov model (data/markov2_enwik8.xz) and say so in the JSON line's `data` field."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/test/enwik8.xz"
dst = os.path.join(ROOT, "data", "enwik8.xz")
if not os.path.exists(src):
    sys.exit(f"{src} not found")
os.makedirs(os.path.dirname(dst), exist_ok=True)
shutil.copyfile(src, dst)
print("copied", src, "->", dst, os.path.getsize(dst), "bytes")
