"""CPU tier: the encoder's CUDA kernel SOURCE executed by the SIMT emulator (tests/simt/) against the oracle.

The emulator compiles orz_b200/csrc/*.cu as host C++ (-DORZ_SIMT_EMU): every CUDA thread is a fibre, warp
collectives and block barriers block until all named lanes arrive, the seqlock / probe-mailbox protocols of stage A
run for real.  It checks the kernel logic bit for bit without a GPU and, with SIMT_SEED, under randomised
interleavings of the warps.  It is test infrastructure: the product loader refuses the emulator build
(test_loader_refuses_the_emulator_build), and the `-m gpu` tests remain the parity tests proper.

The decoder kernel is not emulated: it runs its serial chain redundantly on all 32 lanes of one converged warp
(lock step is part of its design), which a fibre-per-lane model cannot reproduce.
"""
import hashlib
import json
import os
import subprocess
import sys

import pytest

from oracle import pyoracle
from tests import datagen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIMT = os.path.join(ROOT, "tests", "simt")
EMU_LIB = os.path.join(SIMT, "liborz_b200_emu.so")


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-C", SIMT, "liborz_b200_emu.so"], stdout=subprocess.DEVNULL)
    return EMU_LIB


def _run(emu, spec, env_extra=None, timeout=600):
    env = dict(os.environ, ORZ_B200_LIB=emu, ORZ_B200_TEST_EMULATOR="1", SIMT_SWITCH_LIMIT="20000000000")
    env.setdefault("ORZ_B200_V7_CLUSTER", "1")  # one CTA per stream keeps the emulation short; cluster runs are explicit
    env.update(env_extra or {})
    out = subprocess.run([sys.executable, os.path.join(SIMT, "run_case.py"), json.dumps(spec)], env=env, capture_output=True,
                         text=True, timeout=timeout)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


def _want(data, level):
    enc = pyoracle.encode(data, level)
    return {"len": len(enc), "sha256": hashlib.sha256(enc).hexdigest()}


def _data(c):
    return bytes.fromhex(c["literal"]) if "literal" in c else getattr(datagen, c["gen"])(c["n"], seed=c.get("seed", 1))


TINY = [b"", b"a", b"abcd", b"aaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaa", b"the quick brown fox jumps over the lazy dog. " * 30]


@pytest.mark.parametrize("level", [0, 2])
def test_emulated_kernels_match_oracle_tiny(emu, level):
    cases = [{"literal": d.hex()} for d in TINY]
    got = _run(emu, {"level": level, "cases": cases, "batch": True})
    assert got == [_want(d, level) for d in TINY]


@pytest.mark.parametrize("level,case", [(2, {"gen": "text_like", "n": 60000, "seed": 11}),
                                        (1, {"gen": "low_entropy", "n": 50000, "seed": 12}),
                                        (0, {"gen": "random_bytes", "n": 12000, "seed": 13})])
def test_emulated_kernels_match_oracle(emu, level, case):
    got = _run(emu, {"level": level, "cases": [case]})
    assert got == [_want(_data(case), level)]


@pytest.mark.parametrize("seed", [1, 2])
def test_emulated_kernels_under_random_interleavings(emu, seed):
    """Seeded random scheduling of the fibres: the stage A protocol (probe mailbox, per-bucket seqlocks, re-stamping)
    must produce the same bytes under every interleaving."""
    case = {"gen": "text_like", "n": 20000, "seed": 20 + seed}
    got = _run(emu, {"level": 2, "cases": [case]}, {"SIMT_SEED": str(seed)})
    assert got == [_want(_data(case), 2)]


def test_readers_wait_for_records_whose_stamp_came_first():
    """Fault injection (tests/simt/Makefile: -DORZ_V7_TEST_TEAR=1): every builder publishes the stamp of a record BEFORE its words and
    pauses in between — on the B200 a reader was seen to find the stamp of a new record over words of the slot's previous one
    (DESIGN.md 1a-iv).  The decide team and the repairs must notice through the record checksums, wait, and still produce the
    oracle's stream; the library's counters must show that they did notice."""
    lib = os.path.join(SIMT, "liborz_b200_emu_tear.so")
    subprocess.check_call(["make", "-C", SIMT, "liborz_b200_emu_tear.so"], stdout=subprocess.DEVNULL)
    case = {"gen": "text_like", "n": 30000, "seed": 31}
    env = dict(os.environ, ORZ_B200_LIB=lib, ORZ_B200_TEST_EMULATOR="1", SIMT_SWITCH_LIMIT="20000000000", ORZ_B200_V7_CLUSTER="1",
               ORZ_B200_TORN="1", SIMT_SEED="3")
    out = subprocess.run([sys.executable, os.path.join(SIMT, "run_case.py"), json.dumps({"level": 2, "cases": [case]})], env=env,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    assert json.loads(out.stdout.strip().splitlines()[-1]) == [_want(_data(case), 2)]
    seen = [ln for ln in out.stderr.splitlines() if "incomplete records seen" in ln]
    assert seen, out.stderr[-500:]
    team, repairs = (int(x) for x in __import__("re").findall(r"(\d+) by", seen[-1]))
    assert team + repairs > 0, seen[-1]
    # the mutation: the same fault with the readers' checks off must break the stream, or this test proves nothing
    lib2 = os.path.join(SIMT, "liborz_b200_emu_tear_nocheck.so")
    subprocess.check_call(["make", "-C", SIMT, "liborz_b200_emu_tear_nocheck.so"], stdout=subprocess.DEVNULL)
    out2 = subprocess.run([sys.executable, os.path.join(SIMT, "run_case.py"), json.dumps({"level": 2, "cases": [case]})],
                          env=dict(env, ORZ_B200_LIB=lib2), capture_output=True, text=True, timeout=900)
    assert out2.returncode != 0 or json.loads(out2.stdout.strip().splitlines()[-1]) != [_want(_data(case), 2)]


def test_emulated_legacy_v3_variants(emu):
    """v3 (single-thread executor, per-bucket seqlocks; ring-record layout) is the many-stream kernel; pin it for one stream too."""
    case = {"gen": "text_like", "n": 20000, "seed": 31}
    got = _run(emu, {"level": 2, "cases": [case]}, {"ORZ_B200_LEGACY_STAGE_A": "v3s"})
    assert got == [_want(_data(case), 2)]
    got = _run(emu, {"level": 2, "cases": [case]}, {"ORZ_B200_LEGACY_STAGE_A": "v3"})
    assert got == [_want(_data(case), 2)]


def test_emulated_legacy_v4_matches_oracle(emu):
    """Stage A v4 (windowed warp resolver, lock-free builders; parse_v4.cuh): the ring-record layout's one-stream kernel, used for
    LZCfg depths beyond the key lists' 48 candidates."""
    cases = [{"gen": "text_like", "n": 30000, "seed": 51}, {"gen": "low_entropy", "n": 30000, "seed": 52}]
    got = _run(emu, {"level": 2, "cases": cases, "batch": True}, {"ORZ_B200_LEGACY_STAGE_A": "v4"})
    assert got == [_want(_data(c), 2) for c in cases]


@pytest.mark.parametrize("cluster,seed,n", [("1", None, 40000), ("1", "3", 20000), ("8", None, 6000)])
def test_emulated_parse_v7(emu, cluster, seed, n):
    """parse_v7.cuh, the default: key lists in HBM, builders publishing probe records into CTA 0's shared memory (over DSMEM when
    the stream owns a cluster: the emulator runs the CTAs of a cluster together), decide-only resolver, committer warp."""
    env = {"ORZ_B200_V7_CLUSTER": cluster}
    if seed:
        env["SIMT_SEED"] = seed
    cases = [{"gen": "text_like", "n": n, "seed": 71}, {"gen": "low_entropy", "n": n, "seed": 72}]
    got = _run(emu, {"level": 2, "cases": cases, "batch": True}, env)
    assert got == [_want(_data(c), 2) for c in cases]


def test_emulator_flags_divergent_collectives(tmp_path):
    """The runtime fails a launch in which lanes of one warp meet in a collective from different source sites."""
    src = tmp_path / "div.cpp"
    src.write_text('''
#include "simt_rt.h"
#include <cstdio>
static void kern(unsigned *out) {
    unsigned v = threadIdx.x;
    if (threadIdx.x & 1) v = __shfl_sync(0xffffffffu, v, 0);   // odd lanes here ...
    else v = __shfl_sync(0xffffffffu, v, 1);                    // ... even lanes there: undefined on the GPU
    out[threadIdx.x] = v;
}
int main() {
    unsigned out[32];
    simt::launch(dim3(1), dim3(32), 0, [&]() { kern(out); });
    const char *e = simt::last_error();
    printf("%s\\n", e ? e : "no error");
    return e ? 0 : 1;
}
''')
    exe = tmp_path / "div"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", SIMT, "-o", str(exe), str(src), os.path.join(SIMT, "simt_rt.cpp")])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "different source sites" in out.stdout


@pytest.mark.parametrize("env", [{"ORZ_B200_LEGACY_STAGE_A": "v3"}, {"ORZ_B200_LEGACY_STAGE_A": "v4"}, {"ORZ_B200_V7_CLUSTER": "1"},
                                 {"ORZ_B200_V7_CLUSTER": "1", "SIMT_SEED": "6"}])
def test_emulated_long_chain_walks(emu, env):
    """Chains far longer than one gather round (datagen.long_chains): multi-round walks in v3/v4, two-round candidate lists in v7."""
    case = {"gen": "long_chains", "n": 30000, "seed": 4}
    got = _run(emu, {"level": 2, "cases": [case]}, env)
    assert got == [_want(_data(case), 2)]


def test_emulated_operator_interface(emu):
    """orz_lzenc_new/encode/forward (chunk mode, host-driven) through the emulated kernels."""
    case = {"gen": "text_like", "n": 25000, "seed": 54}
    got = _run(emu, {"level": 2, "cases": [case], "operators": True}, {"ORZ_B200_V7_CLUSTER": "1"})
    assert got == [_want(_data(case), 2)]


def test_emulated_parse_v4_under_random_interleavings(emu):
    case = {"gen": "text_like", "n": 20000, "seed": 53}
    got = _run(emu, {"level": 1, "cases": [case]}, {"ORZ_B200_LEGACY_STAGE_A": "v4", "SIMT_SEED": "7"})
    assert got == [_want(_data(case), 1)]


@pytest.fixture(scope="module")
def scale_model():
    """Emulator AND oracle rebuilt with 128 KiB blocks / 3000-item chunks: block slides, `forward`, stale tails and
    chunk cuts (lib.rs:72-86, lz.rs:82-87,128-131) within 150 KB of input."""
    subprocess.check_call(["make", "-C", SIMT, "liborz_b200_emu_small.so", "orz_oracle_small"], stdout=subprocess.DEVNULL)
    cli = os.path.join(SIMT, "emu_cli_small")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", cli, os.path.join(SIMT, "emu_cli.cpp"), "-L" + SIMT,
                           "-lorz_b200_emu_small", "-Wl,-rpath," + SIMT])
    return cli, os.path.join(SIMT, "orz_oracle_small")


@pytest.mark.parametrize("parse,gen,level,seed", [("v3", "text_like", 2, None), ("v4", "text_like", 2, None),
                                                  ("v4", "low_entropy", 1, "5"), ("v3", "low_entropy", 0, "6"),
                                                  ("7", "text_like", 2, None), ("7", "text_like", 1, "8"), ("7", "long_chains", 2, None),
                                                  ("7", "low_entropy", 0, None)])
def test_scale_model_multi_block_streams(scale_model, tmp_path, parse, gen, level, seed):
    cli, oracle = scale_model
    data = getattr(datagen, gen)(150_000 if seed is None else 100_000, seed=61)  # (random scheduling is the slower mode)
    src, want, got = tmp_path / "in", tmp_path / "want.orz", tmp_path / "got.orz"
    src.write_bytes(data)
    subprocess.check_call([oracle, "encode", str(level), str(src), str(want)], stderr=subprocess.DEVNULL)
    env = dict(os.environ, SIMT_SWITCH_LIMIT="60000000000")
    env.update({"ORZ_B200_V7_CLUSTER": "1"} if parse == "7" else {"ORZ_B200_LEGACY_STAGE_A": parse})
    if seed:
        env["SIMT_SEED"] = seed
    subprocess.run([cli, str(level), str(src), str(got)], env=env, check=True, timeout=900, stderr=subprocess.DEVNULL)
    assert got.read_bytes() == want.read_bytes()
    back = tmp_path / "back"
    subprocess.check_call([oracle, "decode", str(want), str(back)], stderr=subprocess.DEVNULL)
    assert back.read_bytes() == data


def test_loader_refuses_the_emulator_build(emu):
    env = dict(os.environ, ORZ_B200_LIB=emu)
    env.pop("ORZ_B200_TEST_EMULATOR", None)
    out = subprocess.run([sys.executable, "-c", "import orz_b200; orz_b200.device_count()"], env=env, cwd=ROOT, capture_output=True,
                         text=True)
    assert out.returncode != 0 and "no CPU fallback" in out.stderr
