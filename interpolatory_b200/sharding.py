"""Work partitioning across GPUs: frame pairs are independent, so pair k goes to
rank k mod N -- no data-path collective.  One process per GPU; torch.distributed is used only
for the barrier and the max-over-ranks reduction of timings."""


def pairs_of_rank(n_pairs, rank, world_size):
    """Pair indices handled by `rank` (round-robin by pair index)."""
    return list(range(rank, n_pairs, world_size))


def runs_of_rank(n_pairs, rank, world_size):
    """Contiguous split alternative: [start, stop) of the pairs of `rank`; every source frame is
    then uploaded to at most two GPUs."""
    base, extra = divmod(n_pairs, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def frames_needed(pairs):
    """Source frame indices a rank must hold for its pairs (pair k needs frames k and k+1)."""
    s = set()
    for k in pairs:
        s.add(k); s.add(k + 1)
    return sorted(s)


def band_split(n_block_rows, n_bands):
    """8K single-pair mode: block rows [start, stop) of each horizontal band."""
    out = []
    for b in range(n_bands):
        start, stop = runs_of_rank(n_block_rows, b, n_bands)
        out.append((start, stop))
    return out
