"""Host-side partition arithmetic of the N > 1 path: which CTU rows of a frame a rank searches, which rows of the
plane it contributes to the all-gather, and how the per-PU records of unequal bands are padded for the gather and put back in
frame order.  Pure integer logic, shared by bench.py and the world_size-2 gloo tests; the compute is the C ABI's band form
(x265b200_me_frame_params.firstCtuRow / ctuRows)."""


def band_rows(ctu_rows, rank, world):
    """(first CTU row, number of CTU rows) of rank's band; bands tile [0, ctu_rows) in rank order and differ by at most one row"""
    lo = (ctu_rows * rank) // world
    hi = (ctu_rows * (rank + 1)) // world
    return lo, hi - lo


def plane_chunk(rows_px, world):
    """rows per rank of the equal-sized all-gather of a plane's picture rows (the last chunk may run past the picture: callers
    allocate world * chunk rows)"""
    return (rows_px + world - 1) // world


def records_per_band(ctu_rows, ctu_cols, pus_per_ctu, num_refs, rank, world):
    """number of {mvx, mvy, cost} records rank produces: [ref][ctu of the band][pu]"""
    return num_refs * band_rows(ctu_rows, rank, world)[1] * ctu_cols * pus_per_ctu


def assemble_bands(parts, ctu_rows, ctu_cols, pus_per_ctu, num_refs, world):
    """parts[r]: rank r's gathered numpy array [>= records_per_band][3] (padded to the largest band).  Returns
    [num_refs][ctu_rows * ctu_cols * pus_per_ctu][3] in the whole-frame order of x265b200_me_frame_ex_dev."""
    import numpy as np
    out = []
    for ref in range(num_refs):
        rows = []
        for r in range(world):
            n = band_rows(ctu_rows, r, world)[1] * ctu_cols * pus_per_ctu
            rows.append(parts[r][ref * n:(ref + 1) * n])
        out.append(np.concatenate(rows, axis=0))
    return np.stack(out)


def triples_of_rank(num_triples, rank, world):
    """lookahead-triple sharding: the (p0, p1, b) cost estimates of a batch are independent
    (slicetype.cpp:1942-1968), so triple i goes to rank i % world; returns this rank's indices in batch order"""
    return list(range(rank, num_triples, world))


def triples_per_rank_max(num_triples, world):
    return (num_triples + world - 1) // world


def assemble_triples(parts, num_triples, world):
    """parts[r]: rank r's gathered numpy array [triples_per_rank_max][...] (padded).  Returns [num_triples][...] in batch order."""
    import numpy as np
    out = np.empty((num_triples,) + tuple(parts[0].shape[1:]), dtype=parts[0].dtype)
    for r in range(world):
        idx = triples_of_rank(num_triples, r, world)
        out[idx] = parts[r][:len(idx)]
    return out
