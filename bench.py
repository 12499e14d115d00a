#!/usr/bin/env python
"""bench.py — BASELINE.json's metric: encode MB/s on enwik8 at -l2, bit-identical output, beside the CPU path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload exact|segments]

Workload (config.workload): BASELINE.json configs[1] — the 100,000,000-byte enwik8 corpus encoded at -l2 as ONE
exact orz stream per GPU (byte-identical to `orz encode -l2 enwik8`).  One orz stream is a serial dependency
chain and does not shard ("replicas only"), so --gpus N runs N independent replicas, one
per rank, followed by one NCCL all-gather of the compressed spans; scaling is "weak".

A step = one pass of the hot path over the whole input.
  value : input resident in HBM, timed with CUDA events on the library's stream (state reset + all kernels).
  e2e   : the public C-ABI call orz_encode_stream() with pinned HOST buffers (allocation, H2D, kernels, D2H inside).
  roofline : the dominant kernel (stage A: orz_parse7_kernel for up to 8 streams per GPU), algorithmic bytes (input read + stream
             written) over its summed CUDA-event launch time, against the measured HBM peak in MEASURED_PEAKS.json.
  decode   : (N = 1) sub-record: the oracle-compressed corpus decoded on the GPU (BASELINE configs[4]), bit-identical flag.
  segments : sub-record measured in the same run: enwik8 x 8 as 16 MiB segments (BASELINE configs[2]/[3] shape), strong scaling over
             the N ranks, spans gathered with one variable-length collective; a sample segment is compared with the oracle.
  cpu_baseline : the C oracle (a port of the reference; no Rust toolchain exists here) on this box's host cores.
--impl reference times that oracle as the reference arm (the real Rust build cannot run on the GPU box).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_BYTES = 100_000_000
LEVEL = 2
METRIC = "encode MB/s on enwik8 at -l2, bit-identical output"


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_encode_threads(data_list, level):
    """Run the oracle on len(data_list) independent streams, one thread each (ctypes releases the GIL)."""
    from oracle import pyoracle
    outs = [None] * len(data_list)

    def work(i):
        outs[i] = pyoracle.encode(data_list[i], level)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(data_list))]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return time.perf_counter() - t0, outs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="exact", choices=["exact", "segments"])
    ap.add_argument("--segment-bytes", type=int, default=1 << 20)
    ap.add_argument("--bytes", type=int, default=N_BYTES, help="debug only: shrink the corpus (prints invalid:true)")
    ap.add_argument("--direction", default="encode", choices=["encode", "decode"],
                    help="decode: BASELINE configs[4], the oracle-compressed corpus decoded on the GPU (extra line, not the headline metric)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = args.gpus
    from tools import textgen
    data, data_desc = textgen.corpus(args.bytes)
    if args.workload == "exact":
        streams = [data]
        wl = f"BASELINE configs[1]: {args.bytes:,} B corpus as one exact orz stream per GPU, -l{LEVEL}"
    else:
        sb = args.segment_bytes
        streams = [data[i:i + sb] for i in range(0, data.size, sb)]
        n_all = len(streams)
        if world > 1:  # independent streams shard across ranks (stream i -> rank i % N); total work is fixed
            from orz_b200.dist import shard_streams
            streams = [streams[i] for i in shard_streams(n_all, rank, world)]
        wl = (f"segment-parallel (extension): {args.bytes:,} B corpus as {n_all} independent orz streams of {sb:,} B "
              f"sharded over {world} GPU(s), -l{LEVEL}")
    config = {"workload": wl, "level": LEVEL, "mode": "exact-stream" if args.workload == "exact" else "segment-parallel",
              "streams_per_gpu": len(streams), "multi_gpu": "replicas only (one orz stream does not shard)",
              "l2": "L2 flushed between timed iterations (256 MiB write)"}
    sharded = args.workload == "segments" and world > 1
    base = {"metric": METRIC if args.direction == "encode" else "decode MB/s on the oracle-compressed corpus (BASELINE configs[4])",
            "unit": "MB/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "u8",
            "data": data_desc if data_desc.startswith("synthetic") else "real: " + data_desc,
            "config": config}
    if args.bytes != N_BYTES:
        base["invalid"] = "debug corpus size"

    # ------------------------------------------------------------------ reference arm (CPU oracle)
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import pyoracle
        pyoracle.build()
        cores = os.cpu_count() or 1
        replicas = min(n_gpus, cores)
        if args.workload == "exact":
            work = [data] * replicas
        else:
            work = streams * 1
        nthreads = min(len(work), cores)
        times = []
        from concurrent.futures import ThreadPoolExecutor
        if args.direction == "decode":  # reference-compressed inputs, timed: the oracle's decoder on all usable host threads
            sizes = [int(np.asarray(w).size) for w in work]
            with ThreadPoolExecutor(nthreads) as ex:
                comp = list(ex.map(lambda s: pyoracle.encode(s, LEVEL), work))
        for it in range(args.warmup + args.steps):
            if args.direction == "decode":
                t0 = time.perf_counter()
                with ThreadPoolExecutor(nthreads) as ex:
                    list(ex.map(lambda cs: pyoracle.decode(cs[0], cs[1] + 64), zip(comp, sizes)))
                dt = time.perf_counter() - t0
            elif args.workload == "exact":
                dt, _ = oracle_encode_threads(work, LEVEL)
            else:  # independent segments over all cores
                t0 = time.perf_counter()
                with ThreadPoolExecutor(nthreads) as ex:
                    list(ex.map(lambda s: pyoracle.encode(s, LEVEL), work))
                dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
        total_bytes = sum(int(np.asarray(w).size) for w in work) * (n_gpus / replicas if args.workload == "exact" else 1)
        sec = sum(times) / len(times)
        v = total_bytes / sec / 1e6
        out = dict(base)
        out.update({"impl": "reference", "value": v, "ms_per_step": sec * 1e3,
                    "cpu_baseline": {"value": v, "unit": "MB/s", "cores": nthreads, "kind": "port",
                                     "sample": f"whole workload per step ({total_bytes:,.0f} B), C oracle port of the "
                                               "reference (no Rust toolchain on this box), %d host cores present" % cores},
                    "e2e": {"value": v, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    "gpu_launches": 0})
        print(json.dumps(out))
        return

    # ------------------------------------------------------------------ our arm (CUDA)
    import torch
    import orz_b200
    from orz_b200 import Batch
    from orz_b200._lib import load, check, LZCfg

    if not torch.cuda.is_available() or orz_b200.device_count() < 1:
        raise SystemExit("bench.py: no sm_100 CUDA device (orz_b200 has no CPU fallback)")
    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    decode_dir = args.direction == "decode"
    if decode_dir:  # the inputs of the hot path are the oracle-compressed streams (reference-compressed, BASELINE configs[4])
        from oracle import pyoracle
        raw_streams = streams
        streams = [np.frombuffer(pyoracle.encode(s, LEVEL), dtype=np.uint8) for s in raw_streams]
        raw_lens = [int(s.size) for s in raw_streams]
    in_lens = [int(s.size) for s in streams]
    if decode_dir:
        batch = Batch(Batch.DECODE, in_lens, out_caps=[r + 64 for r in raw_lens], device=dev)
    else:
        batch = Batch(Batch.ENCODE, in_lens, cfg=LEVEL, device=dev)
    for i, s in enumerate(streams):
        batch.upload(i, s)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{dev}")

    # pinned host buffers for the e2e call
    L = load()
    cfg = LZCfg.from_level(LEVEL)
    pin_in = [torch.from_numpy(np.ascontiguousarray(s)).pin_memory() for s in streams]
    caps = [r + 64 for r in raw_lens] if decode_dir else [L.orz_encode_bound(n) for n in in_lens]
    pin_out = [torch.empty(c, dtype=torch.uint8).pin_memory() for c in caps]
    n = len(streams)
    inp = (C.c_void_p * n)(*[t.data_ptr() for t in pin_in])
    inl = (C.c_size_t * n)(*in_lens)
    outp = (C.c_void_p * n)(*[t.data_ptr() for t in pin_out])
    outc = (C.c_size_t * n)(*caps)
    outl = (C.c_size_t * n)()

    def gather_spans():
        """North-star's final collective: all-gather of the per-stream compressed spans (lengths, then padded spans)."""
        if world <= 1:
            return
        from orz_b200 import dist as odist
        device = torch.device("cuda", dev)
        spans = [odist.device_span(batch.result_dev(i), batch.result_len(i), device) for i in range(n)]
        odist.gather_spans(spans, device)
        torch.cuda.synchronize()

    def measure():
        """W warm-up + K timed passes (resident and end-to-end); returns the per-step lists and the clock sample."""
        dev_ms, parse_ms, emit_ms, parse_launches, launches = [], [], [], [], []
        wall = []
        e2e_s = []
        sampler = None
        for it in range(args.warmup + args.steps):
            flush.fill_(it & 0xFF)
            if it == args.warmup:
                barrier()
                sampler = ClockSampler(dev) if rank == 0 else None
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            batch.run()
            gather_spans()
            t1 = time.perf_counter()
            if it >= args.warmup:
                wall.append(t1 - t0)
                dev_ms.append(batch.timing(4)[0])
                pm, pk = batch.timing(3 if decode_dir else 0)
                em, ek = (0.0, 0) if decode_dir else batch.timing(1)
                bm, bk = (0.0, 0) if decode_dir else batch.timing(2)
                parse_ms.append(pm)
                emit_ms.append(em)
                parse_launches.append(pk)
                launches.append(pk + ek + bk)
            # e2e through the public C ABI with host buffers (one warm-up pass, then every timed step)
            if it < args.warmup - 1:
                continue
            flush.fill_((it + 1) & 0xFF)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if decode_dir:
                check(L.orz_decode_batch(dev, n, inp, inl, outp, outc, outl), "orz_decode_batch")
            else:
                check(L.orz_encode_batch(dev, C.byref(cfg), n, inp, inl, outp, outc, outl), "orz_encode_batch")
            t1 = time.perf_counter()
            if it >= args.warmup:
                e2e_s.append(t1 - t0)
        barrier()
        clocks = sampler.stop() if sampler else None
        return dev_ms, parse_ms, emit_ms, parse_launches, launches, wall, e2e_s, clocks

    dev_ms, parse_ms, emit_ms, parse_launches, launches, wall, e2e_s, clocks = measure()
    comp_bytes = sum(batch.result_len(i) for i in range(n))
    # cross-check e2e output against the resident run
    assert sum(int(outl[i]) for i in range(n)) == comp_bytes
    total_in = sum(in_lens)
    plain_bytes = sum(raw_lens) if decode_dir else total_in   # MB/s are uncompressed bytes per second, like progress.rs:72-83

    # max over ranks of the per-step device time (+ gather for N>1, measured by wall clock around a device sync)
    step_s = (sum(dev_ms) / len(dev_ms)) / 1e3 if world == 1 else sum(wall) / len(wall)
    e2e_step_s = sum(e2e_s) / len(e2e_s)
    plain_all, in_all, out_all = plain_bytes, total_in, comp_bytes
    if world > 1:
        tt = torch.tensor([step_s, e2e_step_s], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        step_s, e2e_step_s = float(tt[0]), float(tt[1])
        ts = torch.tensor([plain_bytes, total_in, comp_bytes], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        plain_all, in_all, out_all = (int(x) for x in ts.tolist())
    # ------------------------------------------------------------------ bit-identity of what was just timed (every rank checks its own stream)
    identical = None
    if args.workload == "exact":
        from oracle import pyoracle
        pyoracle.build()
        sample = raw_streams[0] if decode_dir else streams[0]
        t0 = time.perf_counter()
        want = pyoracle.decode(streams[0], int(np.asarray(sample).size) + 64) if decode_dir else pyoracle.encode(sample, LEVEL)
        cpu_s = time.perf_counter() - t0
        host = pin_out[0].a if hasattr(pin_out[0], "a") else pin_out[0].numpy()  # (.a: the torch stand-in of the CPU flow test)
        identical = batch.download(0) == want and host[: int(outl[0])].tobytes() == want  # resident path AND host-buffer path
        if world > 1:
            ok = torch.tensor([1 if identical else 0], dtype=torch.int32, device=f"cuda:{dev}")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            identical = bool(int(ok[0]))
    batch.close()
    del flush

    # ------------------------------------------------------------------ sub-record: independent 16 MiB segments sharded over the ranks
    seg_rec = None
    if not decode_dir and args.workload == "exact" and args.bytes == N_BYTES and not os.environ.get("ORZ_BENCH_NO_SEGMENTS"):
        seg_rec = segments_subrecord(torch, dist if world > 1 else None, rank, world, dev, np.asarray(data))
    # ------------------------------------------------------------------ sub-record: decode (BASELINE configs[4]), one GPU
    dec_rec = None
    if rank == 0 and world == 1 and not decode_dir and args.workload == "exact" and not os.environ.get("ORZ_BENCH_NO_DECODE"):
        dec_rec = decode_subrecord(np.asarray(streams[0]), want, dev)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = read_peaks()
    p_ms = sum(parse_ms) / len(parse_ms)
    p_launch = sum(parse_launches) / len(parse_launches)
    alg_bytes = total_in + comp_bytes  # encode = N_in read + N_out written
    achieved = alg_bytes / (p_ms / 1e3) / 1e9
    # stage A kernel of this run (mirrors launch_parse in encode_kernels.cu / batch_create in capi.cu)
    if decode_dir:
        kernel = "orz_decode_kernel"
    elif n <= 8 and not os.environ.get("ORZ_B200_LEGACY_STAGE_A"):
        kernel = "orz_parse7_kernel<16,%s>" % os.environ.get("ORZ_B200_V7_CLUSTER", "8")
    else:
        kernel = "orz_parse4_kernel<16,4>" if n == 1 else ("orz_parse_kernel<22,16>" if n <= 296 else "orz_parse_kernel<11,8>")
    # DRAM traffic of the kernel from the committed ncu --set full capture (profiles/traffic.json), scaled from the captured launch to
    # this run's average launch by the algorithmic bytes of the two (the capture is of a shorter launch of the same kernel)
    traffic = traffic_src = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            key = kernel.split("<")[0]
            if key + "_dram_bytes_per_launch" in tj:
                cap_alg = float(tj.get(key + "_algorithmic_bytes_that_launch", 0) or 0)
                traffic = float(tj[key + "_dram_bytes_per_launch"]) * ((alg_bytes / p_launch) / cap_alg if cap_alg > 0 else 1.0)
                traffic_src = tj.get(key + "_source")
        except Exception:
            traffic = traffic_src = None
    out = dict(base)
    out["config"] = dict(config, stage_a=kernel, bit_identical_to_oracle=identical,
                         oracle_pin="two independent restatements agree; NOT pinned against real orz bytes (no Rust toolchain; tools/make_ref_golden.sh)")
    if identical is False:
        out["invalid"] = "the timed stream differs from the oracle's"
    out.update({
        "value": plain_all / step_s / 1e6,
        "ms_per_step": step_s * 1e3,
        "e2e": {"value": plain_all / e2e_step_s / 1e6, "unit": "MB/s", "h2d_bytes_per_step": in_all,
                "d2h_bytes_per_step": out_all},
        "gpu_launches": int(sum(launches)),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes / p_launch, "launches_per_step": p_launch,
                     "avg_launch_ms": p_ms / p_launch,
                     "note": "one orz stream is ONE serial dependency chain (SURVEY 0-2): the kernel is bound by dependent-instruction and "
                             "shared-memory latency of the resolver warp, not by HBM; the fraction is expected to be ~1e-6..1e-5"},
        "kernel_ms_per_step": ({"decode": p_ms} if decode_dir else {"parse": p_ms, "emit": sum(emit_ms) / len(emit_ms)}),
        "compressed_bytes": total_in if decode_dir else comp_bytes,
    })
    if identical is not None:
        out["bit_identical_to_oracle"] = identical
    if n_gpus == 1 and args.workload == "exact":
        nb = int(np.asarray(sample).size)
        out["cpu_baseline"] = {"value": nb / cpu_s / 1e6, "unit": "MB/s", "cores": 1, "kind": "port",
                               "sample": f"the whole {nb:,} B corpus as one stream, one pass, C oracle (port of the reference; Rust cannot "
                                         f"be built here); host has {os.cpu_count()} cores"}
    if seg_rec:
        out["segments"] = seg_rec
    if dec_rec:
        out["decode"] = dec_rec
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


SEG_BYTES = 1 << 24
SEG_COPIES = 8


def segments_subrecord(torch, dist, rank, world, dev, corpus):
    """enwik8 x 8 (800 MB) cut into 16 MiB segments = independent orz streams (fresh LZEncoder each, lib.rs:64), segment i on rank
    i % N, every rank encodes its shard resident in HBM, ONE gather of the spans at the end (orz_b200.dist.gather_spans).  Strong
    scaling: the work is fixed, value = 800 MB / max-over-ranks time.  Rank 0 checks its first and last segment against the oracle."""
    import hashlib
    from orz_b200 import Batch
    from orz_b200 import dist as odist
    from oracle import pyoracle
    total = corpus.size * SEG_COPIES
    nseg = (total + SEG_BYTES - 1) // SEG_BYTES

    def seg(i):  # bytes [i*SEG, (i+1)*SEG) of the corpus repeated SEG_COPIES times
        lo, hi = i * SEG_BYTES, min((i + 1) * SEG_BYTES, total)
        idx = np.arange(lo, hi, dtype=np.int64) % corpus.size
        return corpus[idx]
    mine = odist.shard_streams(nseg, rank, world)
    segs = [seg(i) for i in mine]
    b = Batch(Batch.ENCODE, [int(s.size) for s in segs], cfg=LEVEL, device=dev)
    for i, s in enumerate(segs):
        b.upload(i, s)
    device = torch.device("cuda", dev)
    times = []
    for it in range(2):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        b.run()
        if dist is not None:
            spans = [odist.device_span(b.result_dev(i), b.result_len(i), device) for i in range(len(segs))]
            odist.gather_spans(spans, device)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    sec = times[-1]
    ok = None
    if rank == 0 and segs:
        ok = all(hashlib.sha256(b.download(j)).digest() == hashlib.sha256(pyoracle.encode(segs[j], LEVEL)).digest() for j in {0, len(segs) - 1})
    comp = sum(b.result_len(i) for i in range(len(segs)))
    b.close()
    if dist is not None:
        t = torch.tensor([sec], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t[0])
        cb = torch.tensor([comp], dtype=torch.float64, device=device)
        dist.all_reduce(cb, op=dist.ReduceOp.SUM)
        comp = int(cb[0])
    return {"value": total / sec / 1e6, "unit": "MB/s", "scaling": "strong", "bytes": int(total), "segments": int(nseg), "segment_bytes": SEG_BYTES,
            "level": LEVEL, "ms": sec * 1e3, "compressed_bytes": comp, "sample_identical_to_oracle": ok,
            "what": "BASELINE configs[2]/[3] shape: independent 16 MiB segments sharded i % N, one final gather of the spans; second of two passes, "
                    "inputs resident in HBM, wall clock around a device sync, max over ranks"}


def decode_subrecord(plain, comp, dev):
    """BASELINE configs[4] at its smallest size: the oracle-compressed corpus decoded as one exact stream on one GPU."""
    from orz_b200 import Batch
    comp = np.frombuffer(comp, dtype=np.uint8)
    b = Batch(Batch.DECODE, [int(comp.size)], out_caps=[int(plain.size) + 64], device=dev)
    b.upload(0, comp)
    b.run()
    b.run()
    ms, _ = b.timing(4)
    got = b.download(0)
    b.close()
    return {"value": plain.size / (ms / 1e3) / 1e6, "unit": "MB/s", "ms": ms, "bit_identical": got == plain.tobytes(),
            "what": "decode of the oracle's -l2 stream of the corpus, one stream, second of two passes, CUDA events on the library's stream"}


if __name__ == "__main__":
    main()
