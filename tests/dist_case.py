"""Subprocess body of tests/test_gpu_baseline_configs.py::test_config3: `python dist_case.py WORLD NSEG OUTDIR` spawns WORLD ranks;
NSEG copies of enwik8 are the independent streams, stream i goes to rank i % WORLD (orz_b200.dist.shard_streams), every rank encodes its
shard at -l2 on its GPU, ONE gather of the spans (orz_b200.dist.gather_spans), rank 0 writes them to OUTDIR/spans.npz."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(rank, world, nseg, outdir):
    import numpy as np
    import torch
    import torch.distributed as dist
    from orz_b200 import Batch
    from orz_b200 import dist as odist
    from tests import datagen
    ngpu = torch.cuda.device_count()
    nccl = ngpu >= world
    dev_index = rank if nccl else 0
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl" if nccl else "gloo", rank=rank, world_size=world)
    torch.cuda.set_device(dev_index)
    data = datagen.enwik8()
    mine = odist.shard_streams(nseg, rank, world)
    b = Batch(Batch.ENCODE, [len(data)] * len(mine), cfg=2, device=dev_index)
    for i in range(len(mine)):
        b.upload(i, data)
    b.run()
    dev = torch.device("cuda", dev_index)
    spans = [odist.device_span(b.result_dev(i), b.result_len(i), dev) for i in range(len(mine))]
    if not nccl:
        spans = [s.cpu() for s in spans]
    lengths, gathered = odist.gather_spans(spans, dev if nccl else torch.device("cpu"))
    if rank == 0:
        out = {}
        for r in range(world):
            off = 0
            for j, n in enumerate(lengths[r]):
                out[f"s{odist.shard_streams(nseg, r, world)[j]}"] = gathered[r][off:off + n].cpu().numpy()
                off += n
        np.savez(os.path.join(outdir, "spans.npz"), **out)
    dist.barrier()
    b.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    import torch.multiprocessing as mp
    world, nseg, outdir = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    mp.spawn(worker, args=(world, nseg, outdir), nprocs=world, join=True)
