#!/usr/bin/env python
"""benchmark_tool.py — the reference's benchmark-tool (benchmark-tool/src/main.rs:57-122) for this tree.

Like the reference's tool it runs every codec as a SUBPROCESS reading stdin and writing stdout, three rounds each, encode then decode,
checks the MD5 of the decoded output against the input after every round, keeps the minimum time per direction and prints a markdown
table sorted by compressed size.  Codecs: `orz_b200/bin/orz` (the C++ CLI over the CUDA library; needs a B200) at -l0/-l1/-l2, the CPU
oracle `oracle/orz_oracle` (test infrastructure; here only as the comparison arm the reference's own binary would be), and whichever of
gzip / bzip2 / xz / zstd / brotli exist on PATH with the reference's flags (main.rs:20-36).

Time: the reference reports children's user CPU time (getrusage(RUSAGE_CHILDREN), main.rs:111-122); a GPU codec spends its time on
the device, so this tool prints BOTH wall-clock and children's user CPU seconds.

    python tools/benchmark_tool.py [FILE]          (default: the first 20 MB of data/enwik8.xz; --full for all 100 MB)
"""
import hashlib
import os
import resource
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROUNDS = 3


def child_utime():
    return resource.getrusage(resource.RUSAGE_CHILDREN).ru_utime


def md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def bench(tmp, src, name, enc, dec):
    print(f"start benchmarking {name}...", file=sys.stderr)
    want = md5(src)
    enc_wall, enc_cpu, dec_wall, dec_cpu = [], [], [], []
    eo, do = os.path.join(tmp, "enc_output"), os.path.join(tmp, "dec_output")
    for i in range(ROUNDS):
        for cmd, fin, fout, wall, cpu in ((enc, src, eo, enc_wall, enc_cpu), (dec, eo, do, dec_wall, dec_cpu)):
            with open(fin, "rb") as fi, open(fout, "wb") as fo:
                u0, t0 = child_utime(), time.perf_counter()
                rc = subprocess.run(cmd, stdin=fi, stdout=fo, stderr=subprocess.DEVNULL).returncode
                wall.append(time.perf_counter() - t0)
                cpu.append(child_utime() - u0)
            if rc != 0:
                raise RuntimeError(f"{name}: {' '.join(cmd)}: exit status {rc}")
        if md5(do) != want:
            raise RuntimeError(f"{name}.decode: wrong result")
        print(f" => round {i}: encode {enc_wall[-1]:.3f}s wall / {enc_cpu[-1]:.3f}s cpu, decode {dec_wall[-1]:.3f}s wall / {dec_cpu[-1]:.3f}s cpu",
              file=sys.stderr)
    return os.path.getsize(eo), min(enc_wall), min(enc_cpu), min(dec_wall), min(dec_cpu)


def codecs():
    out = []
    orz = os.path.join(ROOT, "orz_b200", "bin", "orz")
    if os.path.exists(orz):
        for l in (0, 1, 2):
            out.append((f"**orz_b200 -l{l}** (B200)", [orz, "encode", "--silent", f"-l{l}"], [orz, "decode", "--silent"]))
    ora = os.path.join(ROOT, "oracle", "orz_oracle")
    if os.path.exists(ora):
        for l in (0, 2):
            out.append((f"orz oracle -l{l} (CPU port of the reference)", [ora, "encode", str(l), "/dev/stdin", "/dev/stdout"],
                        [ora, "decode", "/dev/stdin", "/dev/stdout"]))
    for name, enc, dec in (("gzip -6", ["gzip", "-6"], ["gzip", "-d"]), ("bzip2 -9", ["bzip2", "-9"], ["bzip2", "-d"]),
                           ("xz -6", ["xz", "-6"], ["xz", "-d"]), ("zstd -19", ["zstd", "-19"], ["zstd", "-d"]),
                           ("brotli -9", ["brotli", "-9"], ["brotli", "-d"])):  # main.rs:20-36
        if shutil.which(enc[0]):
            out.append((name, enc, dec))
    return out


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    with tempfile.TemporaryDirectory() as tmp:
        if args:
            src = args[0]
        else:
            import lzma
            src = os.path.join(tmp, "enwik8")
            n = None if "--full" in sys.argv else 20_000_000
            with lzma.open(os.path.join(ROOT, "data", "enwik8.xz")) as f, open(src, "wb") as o:
                o.write(f.read(n) if n else f.read())
        rows = []
        for name, enc, dec in codecs():
            try:
                size, ew, ec, dw, dc = bench(tmp, src, name, enc, dec)
            except Exception as e:  # noqa: BLE001  (a codec that is not usable here is reported, not fatal: e.g. no GPU)
                print(f"{name}: skipped ({e})", file=sys.stderr)
                continue
            rows.append((size, name, ew, ec, dw, dc))
        rows.sort()
        print(f"benchmark for {os.path.getsize(src):,} bytes ({os.path.basename(src)}), min of {ROUNDS} rounds, MD5-verified round trips\n")
        print("| name | compressed size | encode wall | encode cpu | decode wall | decode cpu |")
        print("|---|---|---|---|---|---|")
        for size, name, ew, ec, dw, dc in rows:
            print(f"| {name} | {size:,} | {ew:.3f}s | {ec:.3f}s | {dw:.3f}s | {dc:.3f}s |")


if __name__ == "__main__":
    main()
