"""Read sharding over the GPUs of one box (replaces `sloika/iterators.py:293-351` `imap_mp`).

The reference parallelises over reads with `multiprocessing.Pool(--jobs)`; results come back
unordered (`bin/basecall_network.py:100-101`).  Here it is one process per GPU (torchrun): reads are
dealt to ranks by length so every rank gets about the same number of samples, each rank basecalls
its own reads with a full replica of the weights, and the `(name, score, path, nsamples)` tuples are
gathered on the host.  No data-path collective exists: reads are independent through forward, decode
and sequence assembly, so NCCL/NVLink carry nothing but the final object gather.
"""


def partition_reads(lengths, world_size):
    """Longest-processing-time deal: sort reads by length (descending) and give each to the rank with
    the least samples so far.  Returns `world_size` lists of read indices (each in ascending order)."""
    assert world_size >= 1
    loads = [0] * world_size
    shards = [[] for _ in range(world_size)]
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += int(lengths[i])
    return [sorted(s) for s in shards]


def shard_batch(nbatch, rank, world_size):
    """Contiguous slice of a batch of `nbatch` equal-length chunks owned by `rank`."""
    base, extra = divmod(nbatch, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_results(local_indices, local_results, nreads, group=None):
    """Host-side gather: every rank passes its read indices and results; all ranks receive the full
    list of `nreads` results in input order (torch.distributed object gather; gloo or nccl)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        out = [None] * nreads
        for i, r in zip(local_indices, local_results):
            out[i] = r
        return out
    pieces = [None] * dist.get_world_size(group)
    dist.all_gather_object(pieces, (list(local_indices), list(local_results)), group=group)
    out = [None] * nreads
    for idx, res in pieces:
        for i, r in zip(idx, res):
            out[i] = r
    return out
