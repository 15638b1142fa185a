"""Host-side sharding of the two places the path partitions across GPUs.

* meshing: chunks are independent given read-only neighbour boundary planes -> rank r meshes the chunks
  `chunk_shard(n, r, world)` (sorted chunk id modulo world size); the shards are put back into batch order with
  `merge_mesh_shards` after an all-gather (every raster rank needs every visible mesh).
* rasterization: rows are independent -> rank r owns the stripe `stripe_of(height, r, world)`, exactly the rows
  `Framebuffer::split_into_stripes(world)` gives slice r (framebuffer.rs:392-431); the disjoint stripes are gathered
  to rank 0 (`gather_stripes`), no depth compositing.

Nothing here computes meshes or pixels; it only moves and reorders them (numpy / torch.distributed, any backend).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


def chunk_shard(n_chunks: int, rank: int, world: int) -> np.ndarray:
    """Chunk ids (positions in the sorted chunk list) meshed by `rank`: k with k % world == rank."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return np.arange(rank, n_chunks, world, dtype=np.int32)


def stripe_of(height: int, rank: int, world: int) -> Tuple[int, int]:
    """(y0, rows) of stripe `rank` of `world` (framebuffer.rs:403-427): ceil(height / world) rows per stripe, the
    last ones shorter or empty (rows == 0: that rank has nothing to draw)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    per = (height + world - 1) // world
    y0 = min(rank * per, height)
    return y0, min(per, height - y0)


def balanced_stripes(band_cost, height: int, world: int, band: int = 8, row_cost: float = 0.0) -> List[Tuple[int, int]]:
    """Contiguous stripes of roughly equal WORK instead of equal height: [(y0, rows)] per rank, covering rows
    0 .. height without gaps, boundaries on multiples of `band` rows.  band_cost[b] is the measured work of rows
    [b * band, (b + 1) * band) (e.g. the triangles binned into that tile row, vx_frame_bin_counts summed over x);
    row_cost is added per row for the clear / write-out every row costs.  Rows are independent in the span
    rasterizer (rasterizer.rs:1401-1427), so any contiguous split renders the same frame as split_into_stripes; this
    one keeps the horizon band from landing on a single GPU.  Greedy prefix split: stripe r ends at the first band
    where the running cost reaches (r + 1) / world of the total; a rank may get no rows when world exceeds the bands."""
    if world <= 0:
        raise ValueError("world must be positive")
    n_bands = (height + band - 1) // band
    cost = np.asarray(band_cost, dtype=np.float64).ravel()
    if cost.size != n_bands:
        raise ValueError(f"band_cost has {cost.size} entries, expected {n_bands}")
    rows_in = np.minimum(band, height - band * np.arange(n_bands)).astype(np.float64)
    cum = np.cumsum(np.maximum(cost, 0.0) + row_cost * rows_in)
    total = float(cum[-1]) if n_bands else 0.0
    cuts = [0]
    for r in range(1, world):
        if total <= 0.0:
            b = min(n_bands, (n_bands * r + world - 1) // world)  # no information: equal bands
        else:
            b = int(np.searchsorted(cum, total * r / world, side="left")) + 1
        cuts.append(min(max(b, cuts[-1]), n_bands))
    cuts.append(n_bands)
    out = []
    for r in range(world):
        y0 = min(cuts[r] * band, height)
        y1 = min(cuts[r + 1] * band, height)
        out.append((y0, y1 - y0))
    return out


def stripe_band_cost(bin_entries, bin_tasks) -> np.ndarray:
    """Per 8-row band: what the band adds to a stripe's frame time, in units of one bin entry.  bin_entries / bin_tasks
    are the (tile rows, tile columns) arrays of vx_frame_bin_counts / vx_frame_bin_tasks of a calibration frame.  Fitted
    on a B200 over 16 stripes of the 1280x720 vd12 frame with eight frames in flight (tools/stripe_probe.py):
    t_stripe = 16.9 us + 1.65e-4 us x entries + 3.4e-5 us x tasks (rms error 1.4 us) -- an entry (a triangle's record load,
    its span setup) weighs about as much as five of the (row, 16-pixel block) tasks it expands to; the constant (cull
    kernel, every mesh's setup unit, clear) is the same for every stripe and does not enter the split."""
    e = np.asarray(bin_entries, dtype=np.float64)
    t = np.asarray(bin_tasks, dtype=np.float64)
    return e.sum(axis=1) + t.sum(axis=1) / 4.9


def merge_mesh_shards(shards: Sequence[Dict[str, np.ndarray]], n_chunks: int, world: Optional[int] = None) -> Dict[str, np.ndarray]:
    """Put per-rank mesh shards (dicts as MeshBatch.download(): quads (Q,3), quad_count, slice_offsets (n,6,33),
    face_aabb (n,6,6), has_mesh; shard r holds the chunks chunk_shard(n_chunks, r, world) in that order) back into
    one batch in chunk order.  Quad streams are concatenated chunk by chunk, quad_base recomputed."""
    world = world or len(shards)
    quad_count = np.zeros(n_chunks, dtype=np.uint32)
    slice_offsets = np.zeros((n_chunks, 6, 33), dtype=np.uint32)
    face_aabb = np.zeros((n_chunks, 6, 6), dtype=np.int32)
    has_mesh = np.zeros(n_chunks, dtype=np.uint8)
    for r, sh in enumerate(shards):
        ids = chunk_shard(n_chunks, r, world)
        if ids.size != sh["quad_count"].shape[0]:
            raise ValueError(f"shard {r} has {sh['quad_count'].shape[0]} chunks, expected {ids.size}")
        quad_count[ids] = sh["quad_count"]
        slice_offsets[ids] = sh["slice_offsets"]
        face_aabb[ids] = sh["face_aabb"]
        has_mesh[ids] = sh["has_mesh"]
    quad_base = np.zeros(n_chunks, dtype=np.uint32)
    if n_chunks:
        quad_base[1:] = np.cumsum(quad_count[:-1], dtype=np.uint64).astype(np.uint32)
    total = int(quad_count.sum(dtype=np.uint64))
    quads = np.zeros((total, 3), dtype=np.uint8)
    for r, sh in enumerate(shards):
        ids = chunk_shard(n_chunks, r, world)
        sq = np.asarray(sh["quads"], dtype=np.uint8).reshape(-1, 3)
        sb = np.asarray(sh["quad_base"], dtype=np.int64)
        sc = np.asarray(sh["quad_count"], dtype=np.int64)
        for j, cid in enumerate(ids.tolist()):
            c = int(sc[j])
            if c:
                quads[int(quad_base[cid]):int(quad_base[cid]) + c] = sq[int(sb[j]):int(sb[j]) + c]
    return {"quads": quads, "quad_base": quad_base, "quad_count": quad_count, "slice_offsets": slice_offsets,
            "face_aabb": face_aabb, "has_mesh": has_mesh}


def all_gather_mesh_shards(local: Dict[str, np.ndarray], n_chunks: int, group=None) -> Dict[str, np.ndarray]:
    """All-gather the ranks' mesh shards (ragged quad streams) and merge them; every rank returns the full batch."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    shards: List[Optional[Dict[str, np.ndarray]]] = [None] * world
    dist.all_gather_object(shards, {k: np.ascontiguousarray(v) for k, v in local.items()}, group=group)
    return merge_mesh_shards(shards, n_chunks, world)  # type: ignore[arg-type]


def gather_stripes(stripe, height: int, width: int, dst: int = 0, group=None, stripes: Optional[Sequence[Tuple[int, int]]] = None):
    """Gather the ranks' disjoint stripes (torch tensors (rows_r, width), any device the backend supports) to `dst`
    and return the composed (height, width) frame there (None elsewhere).  `stripes` = [(y0, rows)] per rank
    (e.g. balanced_stripes); default: the equal split of stripe_of.  Stripes are padded to the tallest one for the
    collective."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    layout = [stripe_of(height, r, world) for r in range(world)] if stripes is None else [(int(y), int(r)) for y, r in stripes]
    if len(layout) != world:
        raise ValueError(f"{len(layout)} stripes for {world} ranks")
    per = max(1, max(r for _, r in layout))
    y0, rows = layout[rank]
    if stripe.shape[0] != rows or stripe.shape[1] != width:
        raise ValueError(f"rank {rank}: stripe is {tuple(stripe.shape)}, expected ({rows}, {width})")
    padded = torch.zeros((per, width), dtype=stripe.dtype, device=stripe.device)
    padded[:rows] = stripe
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    frame = torch.empty((height, width), dtype=stripe.dtype, device=stripe.device)
    for r in range(world):
        ry0, rr = layout[r]
        if rr:
            frame[ry0:ry0 + rr] = bufs[r][:rr]
    return frame
