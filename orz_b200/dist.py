"""Multi-GPU plumbing: one process per GPU (torch.distributed), streams sharded round-robin, no data-path collective
except the final all-gather of the compressed spans (north star).  Works on NCCL (CUDA tensors) and gloo (CPU tensors,
used by the world_size-2 CPU tests)."""
import torch
import torch.distributed as dist


def shard_streams(n_streams: int, rank: int, world: int):
    """Indices of the independent streams this rank encodes (round-robin: stream i -> rank i % world)."""
    return list(range(rank, n_streams, world))


def gather_spans(local_spans, device, group=None):
    """All-gather variable-length byte spans.

    local_spans: list of 1-D uint8 tensors on `device` (this rank's compressed streams, in shard order).
    Returns (lengths, spans): lengths[r] = list of span lengths of rank r, spans[r] = concatenated uint8 tensor of rank r.
    Two collectives: one all_gather of the length table, one all_gather of the spans padded to the longest rank.
    """
    world = dist.get_world_size(group)
    n_local = torch.tensor([len(local_spans)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    max_n = max(int(c.item()) for c in counts)
    lens = torch.zeros(max_n + 1, dtype=torch.int64, device=device)
    for i, s in enumerate(local_spans):
        lens[i] = s.numel()
    lens[max_n] = sum(int(s.numel()) for s in local_spans)
    all_lens = [torch.zeros_like(lens) for _ in range(world)]
    dist.all_gather(all_lens, lens, group=group)
    max_bytes = max(int(t[max_n].item()) for t in all_lens)
    buf = torch.zeros(max(max_bytes, 1), dtype=torch.uint8, device=device)
    off = 0
    for s in local_spans:
        buf[off:off + s.numel()] = s
        off += s.numel()
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    lengths = [[int(x) for x in all_lens[r][: int(counts[r].item())].tolist()] for r in range(world)]
    spans = [out[r][: int(all_lens[r][max_n].item())] for r in range(world)]
    return lengths, spans


def device_span(ptr: int, nbytes: int, device):
    """Zero-copy uint8 view of `nbytes` of device memory at `ptr` (orz_batch_result_dev)."""

    class _W:
        __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}

    if nbytes == 0:
        return torch.empty(0, dtype=torch.uint8, device=device)
    return torch.as_tensor(_W(), device=device)
