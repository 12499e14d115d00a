#!/bin/bash
# Pins the oracle against REAL richox/orz output.  Needs what this repo's containers do not have: nightly Rust + network access to
# crates.io (the crate uses portable_simd / likely_unlikely, src/lib.rs:7-8, and three registry crates, Cargo.toml:16-19).
#
#   tools/make_ref_golden.sh /path/to/richox-orz-checkout        (commit 29414c6)
#
# Builds the reference's own CLI (cargo build --release), encodes every case of tools/ref_golden_cases.py at -l0/-l1/-l2 with the
# UNMODIFIED `orz encode --silent`, and writes tests/golden/ref/<case>.l<level>.json = {case, level, len, sha256, sha256_in, orz_commit}.
# tests/test_oracle.py::test_reference_golden_vectors_when_present then holds oracle/orz_oracle.c to those hashes, and every GPU
# parity test inherits the pin.  Also builds bindings/ffi.rs into the crate (cdylib) when --with-ffi is given, to check that the
# shim compiles against the reference tree (link against orz_b200/liborz_b200.so).
set -euo pipefail
REF=${1:?usage: make_ref_golden.sh /path/to/richox-orz [--with-ffi]}
HERE=$(cd "$(dirname "$0")/.." && pwd)
command -v cargo >/dev/null || { echo "cargo not found: this script must run where nightly Rust is installed" >&2; exit 2; }
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT
( cd "$REF" && cargo +nightly build --release )
ORZ="$REF/target/release/orz"
COMMIT=$(cd "$REF" && git rev-parse --short HEAD 2>/dev/null || echo unknown)
python3 "$HERE/tools/ref_golden_cases.py" "$TMP/in"
mkdir -p "$HERE/tests/golden/ref"
for f in "$TMP"/in/*.bin; do
  name=$(basename "$f" .bin)
  for level in 0 1 2; do
    "$ORZ" encode --silent -l "$level" "$f" "$TMP/out.orz"
    "$ORZ" decode --silent "$TMP/out.orz" "$TMP/back.bin"
    cmp "$f" "$TMP/back.bin"
    python3 - "$name" "$level" "$f" "$TMP/out.orz" "$COMMIT" > "$HERE/tests/golden/ref/$name.l$level.json" <<'PY'
import hashlib, json, sys
name, level, fin, fout, commit = sys.argv[1:6]
a, b = open(fin, "rb").read(), open(fout, "rb").read()
print(json.dumps({"case": name, "level": int(level), "len": len(b), "sha256": hashlib.sha256(b).hexdigest(),
                  "sha256_in": hashlib.sha256(a).hexdigest(), "orz_commit": commit}))
PY
  done
done
if [ "${2:-}" = "--with-ffi" ]; then
  cp "$HERE/bindings/ffi.rs" "$REF/src/ffi.rs"
  sed -i 's|^// pub mod ffi;|pub mod ffi;|' "$REF/src/lib.rs"
  ( cd "$REF" && RUSTFLAGS="-L $HERE/orz_b200 -l orz_b200" cargo +nightly build --release --lib )
  echo "ffi.rs compiled into the reference crate (src/lib.rs:10 re-enabled)"
fi
echo "wrote $(ls "$HERE/tests/golden/ref" | wc -l) reference vectors; now run: python -m pytest tests/test_oracle.py -k reference_golden"
