"""One process per GPU: the two exchange steps of the sharded path, on device memory.

* `StripeCompositor` -- screen-stripe rasterization composed on one GPU WITHOUT a collective.  The reference hands every
  Rayon worker a disjoint `&mut` stripe of one framebuffer (framebuffer.rs:392-431, main.rs:581-597); here the composed
  frame lives in the destination GPU's HBM, the other ranks map it through CUDA IPC and their raster kernels store
  their rows straight over NVLink (`vx_render_frame_into`).  Per frame one 32-bit counter per rank crosses the link
  (`vx_signal_flags` / `vx_wait_flags`), plus the acknowledgement that frees a buffer for re-use (double buffered).
  Stripes can be the reference's equal split (`sharding.stripe_of`) or work-balanced (`sharding.balanced_stripes`).
* `exchange_mesh_shards` -- chunk-sharded meshing: every rank meshes its chunks, packs the shard into one block, ONE
  NCCL all-gather moves the blocks, and `vx_mesh_batch_assemble_shards` rebuilds the full batch on every rank (every
  raster GPU needs every visible mesh, binary_greedy.rs:62-78).

torch.distributed is only the plumbing (handle exchange, the all-gather); all data stays in device memory.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import api
from ._lib import VxShardLayout, VxStripeSync

FLAG_STRIDE_WORDS = 32  # one 128-byte line per flag


class _Cai:
    """Raw device memory as a __cuda_array_interface__ object (torch.as_tensor wraps it without a copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3}


def device_bytes_as_tensor(ptr: int, nbytes: int, device):
    import torch
    return torch.as_tensor(_Cai(ptr, nbytes), device=device)


class StripeCompositor:
    """Composed W x H ARGB frame (and optionally depth) in rank `dst`'s memory, written stripe by stripe by all ranks.

    Protocol for frame k (k = 0, 1, ...; buffer k % n_buffers):
      every rank:  [wait ack >= k - n_buffers + 1]  ->  raster kernel stores its rows into the mapped frame
                   ->  signal arrive[rank] = k + 1
      rank dst:    wait arrive[*] >= k + 1  ->  (consume the frame)  ->  signal ack = k + 1 on every rank
    """

    def __init__(self, ctx: api.Context, width: int, height: int, rank: int, world: int, group=None, dst: int = 0,
                 n_buffers: int = 2, want_depth: bool = False, timeout_us: int = 2_000_000, fused_signal: bool = False):
        import torch.distributed as dist
        self.ctx, self.W, self.H, self.rank, self.world, self.dst, self.group = ctx, int(width), int(height), rank, world, dst, group
        self.n_buffers, self.want_depth, self.timeout_us = int(n_buffers), bool(want_depth), int(timeout_us)
        # Who publishes a rank's arrival word.  Fused into the raster kernel, every CTA has to fence its peer stores at system
        # scope before it counts itself out -- 592 CTAs each holding their SM slot for an NVLink round trip; with several
        # frames in flight that costs more (27 -> 34 us per frame at N = 2) than one 32-thread kernel behind the raster kernel,
        # whose completion flushes the stores anyway.  The composing GPU's own stripe stays fused (its stores are local).
        self.fused_signal = bool(fused_signal)
        lib, h = ctx.lib, ctx.handle
        self.frame_bytes = self.W * self.H * 4
        planes = 2 if want_depth else 1
        self._own: List[C.c_void_p] = []
        self._mapped: List[C.c_void_p] = []

        def alloc(nbytes):
            p = C.c_void_p()
            ctx.check(lib.vx_device_alloc(h, nbytes, C.byref(p)))
            self._own.append(p)
            return p

        def export(p):
            buf = (C.c_uint8 * 64)()
            ctx.check(lib.vx_ipc_export(h, p, buf))
            return bytes(buf)

        def open_(handle: bytes):
            p = C.c_void_p()
            ctx.check(lib.vx_ipc_open(h, (C.c_uint8 * 64).from_buffer_copy(handle), C.byref(p)))
            self._mapped.append(p)
            return p

        # every rank owns one acknowledgement word; dst owns the frames and one arrival word per rank
        self.ack_local = alloc(128)
        mine = {"ack": export(self.ack_local)}
        if rank == dst:
            self.frames_local = alloc(self.frame_bytes * planes * self.n_buffers)
            self.arrive_local = alloc(4 * FLAG_STRIDE_WORDS * world)
            mine["frames"] = export(self.frames_local)
            mine["arrive"] = export(self.arrive_local)
        everyone: List[Optional[dict]] = [None] * world
        dist.all_gather_object(everyone, mine, group=group)
        if rank == dst:
            self.frames = self.frames_local.value
            self.arrive = self.arrive_local.value
            self.acks = [self.ack_local.value if r == rank else open_(everyone[r]["ack"]).value for r in range(world)]
        else:
            self.frames = open_(everyone[dst]["frames"]).value
            self.arrive = open_(everyone[dst]["arrive"]).value
            self.acks = None
        self.stripes: List[Tuple[int, int]] = []
        self._call_state = None
        self.set_stripes(None)
        dist.barrier(group=group)

    # ---- layout ---------------------------------------------------------------------------------------------
    def set_stripes(self, stripes: Optional[Sequence[Tuple[int, int]]]):
        """[(y0, rows)] per rank: contiguous, disjoint, covering rows 0 .. H.  None: split_into_stripes(world)."""
        from .sharding import stripe_of
        layout = [stripe_of(self.H, r, self.world) for r in range(self.world)] if stripes is None else [(int(a), int(b)) for a, b in stripes]
        if len(layout) != self.world:
            raise ValueError(f"{len(layout)} stripes for {self.world} ranks")
        y = 0
        for y0, rows in layout:
            if rows < 0 or (rows > 0 and y0 != y):
                raise ValueError(f"stripes must tile the frame top to bottom: {layout}")
            y += rows
        if y != self.H:
            raise ValueError(f"stripes cover {y} of {self.H} rows")
        self.stripes = layout

    def color_ptr(self, frame_no: int) -> int:
        planes = 2 if self.want_depth else 1
        return self.frames + (frame_no % self.n_buffers) * self.frame_bytes * planes

    def depth_ptr(self, frame_no: int) -> int:
        return self.color_ptr(frame_no) + self.frame_bytes if self.want_depth else 0

    # ---- per frame --------------------------------------------------------------------------------------------
    def render(self, batch: api.MeshBatch, view_proj, camera_position, cfg, view_distance: int, frame_no: int,
               compose_release: Optional[int] = None):
        """Enqueue this rank's stripe of frame `frame_no` (asynchronous; cfg is copied with the stripe filled in).
        compose_release (dst only, needs a stripe with rows): fold the per-frame bookkeeping of the composing GPU into the
        raster kernel as well -- its last CTA waits for every rank's arrival word of this frame and then hands the buffer of
        frame `compose_release` back (use frame_no when nothing reads the frame on the host, an older frame otherwise).
        Returns True when that was done (else call complete() / release() or complete_and_release())."""
        ctx, lib, h = self.ctx, self.ctx.lib, self.ctx.handle
        y0, rows = self.stripes[self.rank]
        arrive = self.arrive + 4 * FLAG_STRIDE_WORDS * self.rank
        wait_needed = frame_no >= self.n_buffers  # the buffer is free once dst has consumed frame_no - n_buffers
        if rows > 0:
            # the hand-off rides in the frame's own kernels: the setup kernel waits for the acknowledgement, the arrival word is
            # published behind the raster kernel (or by its last CTA: fused_signal), the composing GPU's raster kernel waits
            # for all arrivals and hands an older buffer back -- one launch graph per frame and rank, no collective.
            # (The ctypes objects are kept between calls: at tens of thousands of frames per second the host side of a frame
            # has to stay in the single-digit microseconds.)
            st = self._call_state
            if st is None:
                st = self._call_state = {"cfg": api.VxFrameConfig(), "sync": VxStripeSync(), "vp": np.zeros(16, np.float32), "cam": np.zeros(3, np.float32),
                                         "rel": (C.c_void_p * self.world)(*self.acks) if self.acks is not None else None}
                st["cfg_ref"], st["sync_ref"] = C.byref(st["cfg"]), C.byref(st["sync"])
                st["vp_p"], st["cam_p"] = st["vp"].ctypes.data_as(C.c_void_p), st["cam"].ctypes.data_as(C.c_void_p)
                st["rel_p"] = C.cast(st["rel"], C.c_void_p) if st["rel"] is not None else None
                st["fn"] = lib.vx_render_frame_stripe
            c, sync = st["cfg"], st["sync"]
            C.memmove(st["cfg_ref"], C.byref(cfg), C.sizeof(api.VxFrameConfig))
            c.stripe_y0, c.stripe_rows = y0, rows
            off = y0 * self.W * 4
            sync.d_wait_flag = self.ack_local.value if wait_needed else None
            sync.wait_value = (frame_no - self.n_buffers + 1) & 0xFFFFFFFF if wait_needed else 0
            sync.signal_after = 1 if (not self.fused_signal and self.rank != self.dst) else 0
            sync.d_signal_flag = arrive
            sync.signal_value = (frame_no + 1) & 0xFFFFFFFF
            sync.timeout_us = self.timeout_us
            fused = compose_release is not None and self.rank == self.dst
            if fused:
                sync.n_arrive, sync.arrive_stride_words, sync.d_arrive_flags = self.world, FLAG_STRIDE_WORDS, self.arrive
                sync.arrive_value = (frame_no + 1) & 0xFFFFFFFF
                sync.release_value = (compose_release + 1) & 0xFFFFFFFF if compose_release >= 0 else 0
                sync.n_release = self.world if compose_release >= 0 else 0
                sync.release_flags = st["rel_p"]
            else:
                sync.n_arrive = 0
                sync.n_release = 0
            if view_proj is not st.get("vp_src") or camera_position is not st.get("cam_src"):
                st["vp"][:] = np.asarray(view_proj, dtype=np.float32).reshape(16)
                st["cam"][:] = camera_position
                # immutable inputs (tuples) may be recognised by identity next time; arrays may be changed in place by the caller
                st["vp_src"] = view_proj if isinstance(view_proj, tuple) else None
                st["cam_src"] = camera_position if isinstance(camera_position, tuple) else None
            rc = st["fn"](h, batch.handle, None, -1, st["vp_p"], st["cam_p"], int(view_distance), st["cfg_ref"],
                          self.color_ptr(frame_no) + off, (self.depth_ptr(frame_no) + off) if self.want_depth else None, st["sync_ref"])
            if rc != 0:
                ctx.check(rc)
            return fused
        # a rank without rows only reports in
        if wait_needed:
            ctx.check(lib.vx_wait_flags(h, self.ack_local, 1, 1, frame_no - self.n_buffers + 1, self.timeout_us))
        flag = (C.c_void_p * 1)(arrive)
        ctx.check(lib.vx_signal_flags(h, flag, 1, frame_no + 1))
        return False

    def complete(self, frame_no: int):
        """dst only: enqueue the wait for every rank's stripe of `frame_no`; later work on the stream sees the frame."""
        assert self.rank == self.dst
        self.ctx.check(self.ctx.lib.vx_wait_flags(self.ctx.handle, C.c_void_p(self.arrive), self.world, FLAG_STRIDE_WORDS, frame_no + 1, self.timeout_us))

    def release(self, frame_no: int):
        """dst only: enqueue the acknowledgement that frame `frame_no` has been consumed (its buffer may be re-used)."""
        assert self.rank == self.dst
        flags = (C.c_void_p * self.world)(*self.acks)
        self.ctx.check(self.ctx.lib.vx_signal_flags(self.ctx.handle, flags, self.world, frame_no + 1))

    def complete_and_release(self, frame_no: int, release_frame_no: Optional[int] = None):
        """dst only: one kernel that waits for every stripe of `frame_no` and then hands the buffer of `release_frame_no`
        (default: the same frame) back to all ranks."""
        assert self.rank == self.dst
        rel = frame_no if release_frame_no is None else release_frame_no
        flags = (C.c_void_p * self.world)(*self.acks)
        self.ctx.check(self.ctx.lib.vx_wait_then_signal(self.ctx.handle, C.c_void_p(self.arrive), self.world, FLAG_STRIDE_WORDS, (frame_no + 1) & 0xFFFFFFFF,
                                                        flags, self.world, (rel + 1) & 0xFFFFFFFF, self.timeout_us))

    def check(self):
        """Synchronise and raise if a wait timed out."""
        t = C.c_int32(0)
        self.ctx.check(self.ctx.lib.vx_wait_status(self.ctx.handle, C.byref(t)))

    def frame_tensor(self, frame_no: int, device):
        """dst only: the composed colour plane of `frame_no` as an (H, W) int32 torch tensor (no copy)."""
        import torch
        assert self.rank == self.dst
        return device_bytes_as_tensor(self.color_ptr(frame_no), self.frame_bytes, device).view(torch.int32).view(self.H, self.W)

    def depth_tensor(self, frame_no: int, device):
        import torch
        assert self.rank == self.dst and self.want_depth
        return device_bytes_as_tensor(self.depth_ptr(frame_no), self.frame_bytes, device).view(torch.float32).view(self.H, self.W)

    def close(self):
        import torch.distributed as dist
        self.ctx.synchronize()
        dist.barrier(group=self.group)  # nobody unmaps / frees while a peer may still store
        for p in self._mapped:
            self.ctx.lib.vx_ipc_close(self.ctx.handle, p)
        self._mapped = []
        dist.barrier(group=self.group)
        for p in self._own:
            self.ctx.lib.vx_device_free(self.ctx.handle, p)
        self._own = []


class MeshShardExchange:
    """Chunk-sharded meshing with the exchange on the device: mesh this rank's chunks, all-gather the packed shards
    (one NCCL collective), assemble the full batch.  Buffers are kept between sweeps."""

    def __init__(self, ctx: api.Context, n_chunks: int, rank: int, world: int, device, group=None):
        from .sharding import chunk_shard
        import torch
        self.ctx, self.n, self.rank, self.world, self.device, self.group = ctx, int(n_chunks), rank, world, device, group
        self.ids = chunk_shard(self.n, rank, world)
        self.d_ids = torch.from_numpy(self.ids).to(device)
        self.rows = (self.n + world - 1) // world
        self.shard: Optional[api.MeshBatch] = None
        self.full: Optional[api.MeshBatch] = None
        self.layout = VxShardLayout()
        self.block = None
        self.blocks = None
        self.totals = torch.zeros(world, dtype=torch.int64, device=device)
        self.mine = torch.zeros(1, dtype=torch.int64, device=device)
        self.stream = torch.cuda.ExternalStream(ctx.stream, device=device)

    def sweep(self, d_voxels: int, d_positions: int, d_neighbors: int, d_uniform_flags: int = 0, host_sync: Optional[bool] = None) -> api.MeshBatch:
        """Re-mesh this rank's shard and rebuild the full batch on every rank.  Returns the full batch.
        The first sweep reads the ragged shard sizes back to size the exchange blocks (two host round trips); later sweeps
        (host_sync=None / False) run without one: mesh kernel -> pack (the block carries the shard's quad total) -> one
        NCCL all-gather -> assemble from the totals on the device.  A shard that outgrows its block makes the next
        .info() of the full batch fail with VX_ERR_CAPACITY: call sweep(..., host_sync=True) once to size the blocks anew
        (check() does that check for you)."""
        import torch
        import torch.distributed as dist
        ctx, lib = self.ctx, self.ctx.lib
        self.shard = api.BinaryGreedyMesher.mesh_batch_subset(d_voxels, d_positions, d_neighbors, d_uniform_flags, self.n, self.d_ids.data_ptr(),
                                                              int(self.ids.size), ctx, batch=self.shard)
        if host_sync is None:
            host_sync = self.block is None or self.full is None
        if not host_sync:
            ctx.check(lib.vx_mesh_shard_pack_async(ctx.handle, self.shard.handle, C.byref(self.layout), C.c_void_p(self.block.data_ptr())))
            with torch.cuda.stream(self.stream):
                self._all_gather(self.blocks, self.block)  # NCCL over NVLink
            h = C.c_void_p(self.full.handle.value)
            ctx.check(lib.vx_mesh_batch_assemble_shards_async(ctx.handle, self.n, self.world, C.c_void_p(self.blocks.data_ptr()), C.byref(self.layout),
                                                              C.c_void_p(d_positions) if d_positions else None, C.byref(h)))
            self.full._host = None
            self.exchanged_bytes = int(self.layout.rank_stride) * self.world
            return self.full
        my_quads = int(self.shard.info().total_quads)  # synchronises: the ragged sizes have to be known to size the exchange
        with torch.cuda.stream(self.stream):
            self.mine.fill_(my_quads)
            self._all_gather(self.totals, self.mine)
            totals = self.totals.cpu().numpy().astype(np.int64)
        need = int(totals.max())
        if self.block is None or need > int(self.layout.quads_capacity):
            cap = need + need // 8 + 1024  # every rank derives the same capacity from the same totals
            ctx.check(lib.vx_shard_layout(self.rows, cap, C.byref(self.layout)))
            self.block = torch.empty(int(self.layout.rank_stride), dtype=torch.uint8, device=self.device)
            self.blocks = torch.empty(int(self.layout.rank_stride) * self.world, dtype=torch.uint8, device=self.device)
        ctx.check(lib.vx_mesh_shard_pack(ctx.handle, self.shard.handle, C.byref(self.layout), C.c_void_p(self.block.data_ptr())))
        with torch.cuda.stream(self.stream):
            self._all_gather(self.blocks, self.block)  # NCCL over NVLink
        h = C.c_void_p(self.full.handle.value if self.full is not None else None)
        ctx.check(lib.vx_mesh_batch_assemble_shards(ctx.handle, self.n, self.world, C.c_void_p(self.blocks.data_ptr()), C.byref(self.layout),
                                                    totals.ctypes.data_as(C.c_void_p), C.c_void_p(d_positions) if d_positions else None, C.byref(h)))
        if self.full is None:
            self.full = api.MeshBatch(ctx, h)
        self.full._host = None
        self.exchanged_bytes = int(self.layout.rank_stride) * self.world
        return self.full

    def check(self) -> bool:
        """After sweeps without host round trips: True if every shard fitted its block (synchronises)."""
        try:
            self.full.info()
            return True
        except api.VxError:
            return False

    def _all_gather(self, out, inp):
        import torch
        import torch.distributed as dist
        if dist.get_backend(self.group) == "nccl":
            dist.all_gather_into_tensor(out, inp, group=self.group)
            return
        # Ranks that share one GPU (the single-GPU CI box runs the two-rank test that way) cannot form an NCCL communicator;
        # the gloo group moves the same bytes through host memory.  Not a production transport.
        self.stream.synchronize()
        parts = [torch.empty(inp.shape, dtype=inp.dtype) for _ in range(self.world)]
        dist.all_gather(parts, inp.cpu(), group=self.group)
        out.copy_(torch.cat([x.reshape(-1) for x in parts]).view(out.dtype).reshape(out.shape))

    def close(self):
        for b in (self.shard, self.full):
            if b is not None:
                b.release()
        self.shard = self.full = None
