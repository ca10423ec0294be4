"""Sharding of reads over GPUs / ranks.  Reads are independent (TranscriptClean.py:365-367) and
the reference itself cuts them into contiguous input-order chunks of ceil(n/k) reads
(split_input, TC.py:544-557); the same rule is used here so that concatenating the shards in
rank order reproduces the `--threads 1` output order.  No collective on the data path."""
from math import ceil


def shard_bounds(n, world):
    """[(lo, hi)] per rank; trailing ranks may be empty (like split_input, which yields fewer chunks)."""
    if world <= 0:
        raise ValueError("world must be positive")
    size = ceil(n / world) if n else 0
    out = []
    for r in range(world):
        lo = min(n, r * size)
        out.append((lo, min(n, lo + size)))
    return out


def concat_texts(parts):
    """Concatenate per-shard {"sam","fa","log","te"} dicts in shard order."""
    out = {"sam": "", "fa": "", "log": "", "te": ""}
    for p in parts:
        for k in out:
            out[k] += p[k]
    return out


def correct_sharded_threads(lines, options, refs_list, worker):
    """One host thread per GPU context (`refs_list`), contiguous shards, ordered concatenation.
    `worker(lines, options, refs)` returns the four texts of one shard."""
    import threading
    bounds = shard_bounds(len(lines), len(refs_list))
    parts = [None] * len(refs_list)
    errs = []

    def run(k):
        try:
            lo, hi = bounds[k]
            parts[k] = worker(lines[lo:hi], options, refs_list[k])
        except Exception as e:      # surfaced after join
            errs.append(e)

    ts = [threading.Thread(target=run, args=(k,)) for k in range(len(refs_list))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errs:
        raise errs[0]
    return concat_texts(parts)


def correct_sharded_dist(lines, options, refs, worker, dist):
    """torch.distributed flavour: every rank corrects its shard; rank 0 receives the texts in rank
    order (gather_object is host-side plumbing, not a data-path collective)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = shard_bounds(len(lines), world)[rank]
    mine = worker(lines[lo:hi], options, refs)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    return concat_texts(gathered) if rank == 0 else None
