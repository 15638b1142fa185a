#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 voxel frame path.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): one full 1280x720 frame at view
distance 12 -- chunk cull (filter A) + AABB reject / draw order (filter B) + project / near-clip / backface cull of every
quad + span rasterization with depth buffer and textured shading -- of the seeded synthetic terrain world (7,153
lattice chunks, the Varied ones meshed with their neighbours), camera (0,10,20) looking down -Z, meshes cached on the
device exactly as the reference caches them between frames (main.rs:225-280).  A "step" is one frame.

One JSON line on stdout (rank 0).  `value` = device-resident frames/s (CUDA events on the launching stream, L2 flushed
between timed frames); `e2e` = the same frame through the public host API (VP + camera uploaded, ARGB frame read back
into pinned host memory, every step); `roofline` describes the dominant kernel; `cpu_baseline` is the C restatement of
the reference CPU path timed on this box's host cores; `extra` carries the second metric of BASELINE.json (chunks
meshed per second) and the per-kernel split.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W, H, VD = 1280, 720, 12
METRIC = "frames_per_sec_1280x720_vd12"
UNIT = "frames/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def ncu_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of `kernel` per launch from the newest committed ncu capture
    (profiles/*_traffic.json, written from an `ncu --set full` run of tools/ncu_target.py), or None."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
    if not files:
        return None
    try:
        return float(json.load(open(files[-1]))[kernel]["traffic"])
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append(line.strip())
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_scene():
    import vx_scenes
    pos, world, p, v, nb = vx_scenes.terrain_scene(VD)
    cam = vx_scenes.main_camera(W, H)
    return pos, world, p, v, nb, cam


# ------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the path (C restatement, see oracle/vx_oracle.h), all host threads
# ------------------------------------------------------------------------------------------------------------------
def cpu_frame_baseline(p, v, nb, cam, min_seconds: float, threads: int, steps=None, warmup=1):
    from oracle import binding as ob
    ref = ob.mesh_chunks(v, nb, None, p)
    vp = cam.view_projection()
    vis = ob.cull_chunks(p, vp, cam.position, VD)
    cfg = ob.default_frame_config(W, H, n_threads=threads)
    atlas = ob.default_atlas()
    has = ref.has_mesh != 0

    def one_frame():
        vis = ob.cull_chunks(p, vp, cam.position, VD)           # filter A
        ids = np.flatnonzero((vis != 0) & has).astype(np.int32)
        return ob.render_frame(ref, ids, vp, cam.position, cfg, atlas)  # filter B + sort + raster

    for _ in range(max(1, warmup)):
        one_frame()
    n, t0 = 0, time.perf_counter()
    while True:
        one_frame()
        n += 1
        el = time.perf_counter() - t0
        if steps is not None:
            if n >= steps:
                break
        elif el >= min_seconds:
            break
    return n / el, n, el


def cpu_mesh_baseline(v, nb, min_seconds: float):
    from oracle import binding as ob
    n, t0 = 0, time.perf_counter()
    while True:
        ob.mesh_chunks(v, nb)
        n += v.shape[0]
        el = time.perf_counter() - t0
        if el >= min_seconds:
            break
    return n / el, n, el


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    pos, world, p, v, nb, cam = build_scene()
    threads = os.cpu_count() or 1
    fps, n, el = cpu_frame_baseline(p, v, nb, cam, 0.0, threads, steps=max(1, args.steps), warmup=max(1, args.warmup))
    out = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": n, "warmup": args.warmup,
        "ms_per_step": 1000.0 * el / n, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": "full frame 1280x720 view distance 12 (filter A + filter B/sort + project/clip/cull + span raster), "
                                                     "meshes cached; BASELINE.json configs[2]",
                                         "chunks": int(pos.shape[0]), "varied_chunks": int(p.shape[0]),
                                         "camera": "(0,10,20) yaw 0 pitch 0 fov 70", "parallelism": f"{threads} host threads (stripes)"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n} full frames; C restatement of the reference CPU path (oracle/), stripe-parallel over all host threads; "
                                   "the Rust reference itself cannot be built in this image (no cargo/rustc)"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_json_line(out)
    return 0


# ------------------------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------------------------
def finish_distributed(dist):
    """All ranks leave together: barrier, tear the process group down, and exit without running interpreter-exit
    destructors (torch's NCCL watchdog otherwise races the CUDA context teardown and aborts the process)."""
    try:
        dist.barrier()
        dist.destroy_process_group()
    finally:
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def run_cuda(args):
    import torch
    from differential_projection_voxel_renderer_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world_size > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=180))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    pos, world, p, v, nb, cam = build_scene()
    vp = cam.view_projection()
    ctx = api.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    # ---- inputs resident in HBM --------------------------------------------------------------------------------
    d_vox = torch.from_numpy(v).to(dev)
    d_nb = torch.from_numpy(nb).to(dev)
    d_pos = torch.from_numpy(p).to(dev)
    n_chunks = int(p.shape[0])

    # chunk-sharded meshing (sorted chunk id modulo n_gpu); the frame needs every visible mesh on every raster GPU
    h = C.c_void_p()
    ctx.check(ctx.lib.vx_mesh_chunks_device(ctx.handle, C.c_void_p(d_vox.data_ptr()), C.c_void_p(d_pos.data_ptr()),
                                            C.c_void_p(d_nb.data_ptr()), None, n_chunks, C.byref(h)))
    batch = api.MeshBatch(ctx, h)
    info = batch.info()
    total_quads = int(info.total_quads)

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def flush_l2():
        with torch.cuda.stream(stream):
            flush.fill_(1)

    # stripe of this rank (framebuffer.rs:392-431: ceil(H / n) rows per stripe)
    cfg = api.default_frame_config(W, H)
    rows_per = (H + world_size - 1) // world_size
    if world_size > 1:
        from differential_projection_voxel_renderer_b200.sharding import stripe_of
        cfg.stripe_y0, cfg.stripe_rows = stripe_of(H, rank, world_size)  # framebuffer.rs:403-427
    cfg_async = api.VxFrameConfig.from_buffer_copy(cfg)
    cfg_async.async_submit = 1

    gather_bufs = None
    if world_size > 1:
        gather_bufs = [torch.empty((rows_per, W), dtype=torch.int32, device=dev) for _ in range(world_size)] if rank == 0 else None
        my_stripe = torch.empty((rows_per, W), dtype=torch.int32, device=dev)

    class _Cai:
        def __init__(self, ptr, shape, typestr):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3}

    def step_device():
        if world_size == 1:
            api.render_frame_device(batch, vp, cam.position, cfg_async, VD, ctx)
        else:  # the raster kernel writes this rank's rows straight into the gather buffer
            api.render_frame_into(batch, vp, cam.position, cfg_async, VD, my_stripe.data_ptr(), 0, ctx)
            with torch.cuda.stream(stream):
                dist.gather(my_stripe, gather_bufs, dst=0)  # composite: disjoint stripes into GPU0

    # ---- warm-up + correctness guard ----------------------------------------------------------------------------
    api.render_frame_device(batch, vp, cam.position, cfg, VD, ctx)  # one synchronous frame sizes the frame scratch
    for _ in range(max(3, args.warmup)):
        step_device()
    ctx.synchronize()
    st = api.frame_stats(ctx)
    launches_per_frame = st.n_kernel_launches

    # ---- timed region: exactly K frames, CUDA events on the launching stream, L2 flushed between frames ---------
    sampler = ClockSampler(local_rank)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    K = max(1, args.steps)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    l0 = ctx.launch_count
    for i in range(K):
        flush_l2()
        starts[i].record(stream)
        step_device()
        ends[i].record(stream)
    torch.cuda.synchronize()
    l1 = ctx.launch_count
    # keep the same load running until the sampler has seen >= 1.5 s of it (the timed frames are ~tens of microseconds)
    t_probe = time.perf_counter()
    while time.perf_counter() - t_probe < 1.5:
        for _ in range(50):  # this rank's frames only: a time-based loop must not contain collectives (ranks would disagree on the count)
            api.render_frame_device(batch, vp, cam.position, cfg_async, VD, ctx)
        ctx.synchronize()
    clocks = sampler.stop()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    total_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / K
    fps = 1000.0 / ms_per_step
    api.frame_stats(ctx)  # surfaces an overflow of an async frame, if any

    # ---- chunk-sharded remesh sweep (BASELINE cfg 4): rank r re-meshes the chunks k with k % N == r; the world's voxel
    #      array is replicated, so halos need no exchange and there is no collective on the path (weak in the sense
    #      that more GPUs are for more chunks; here the world is fixed, so the line is total chunks / max time)
    from differential_projection_voxel_renderer_b200 import sharding
    sub_ids = sharding.chunk_shard(n_chunks, rank, world_size)
    d_sub = torch.from_numpy(sub_ids).to(dev)
    sub_batch = api.BinaryGreedyMesher.mesh_batch_subset(d_vox.data_ptr(), d_pos.data_ptr(), d_nb.data_ptr(), 0, n_chunks,
                                                         d_sub.data_ptr(), int(sub_ids.size), ctx)

    def remesh_shard():
        api.BinaryGreedyMesher.mesh_batch_subset(d_vox.data_ptr(), d_pos.data_ptr(), d_nb.data_ptr(), 0, n_chunks,
                                                 d_sub.data_ptr(), int(sub_ids.size), ctx, batch=sub_batch)

    for _ in range(3):
        remesh_shard()
    ctx.synchronize()
    if dist is not None:
        dist.barrier()
    sh_ms = []
    for _ in range(20):
        flush_l2()
        a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        remesh_shard()
        b2.record(stream)
        torch.cuda.synchronize()
        sh_ms.append(a.elapsed_time(b2))
    shard_ms = float(np.mean(sh_ms))
    if dist is not None:
        t = torch.tensor([shard_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        shard_ms = float(t.item())
    sharded_chunks_per_s = n_chunks / (shard_ms * 1e-3)
    sub_batch.release()

    # ---- the other way to use N GPUs for this metric: every rank renders whole frames (alternate-frame rendering, no
    #      collective, weak scaling in frames).  Reported next to the stripe-sharded headline, not instead of it.
    afr_fps = None
    if world_size > 1:
        cfg_full = api.default_frame_config(W, H)
        api.render_frame_device(batch, vp, cam.position, cfg_full, VD, ctx)  # sizes the scratch for the full frame
        cfg_full_async = api.VxFrameConfig.from_buffer_copy(cfg_full)
        cfg_full_async.async_submit = 1
        for _ in range(3):
            api.render_frame_device(batch, vp, cam.position, cfg_full_async, VD, ctx)
        ctx.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        a_s = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        a_e = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        for i in range(K):
            flush_l2()
            a_s[i].record(stream)
            api.render_frame_device(batch, vp, cam.position, cfg_full_async, VD, ctx)
            a_e[i].record(stream)
        torch.cuda.synchronize()
        t_afr = torch.tensor([sum(x.elapsed_time(y) for x, y in zip(a_s, a_e))], dtype=torch.float64, device=dev)
        dist.all_reduce(t_afr, op=dist.ReduceOp.MAX)
        afr_fps = world_size * K / (float(t_afr.item()) * 1e-3)
        api.frame_stats(ctx)
        api.render_frame_device(batch, vp, cam.position, cfg, VD, ctx)  # back to this rank's stripe

    # ---- e2e at N > 1: every rank renders its stripe through the device API (VP + camera + config uploaded per call),
    #      the stripes are gathered to GPU0 over NVLink and rank 0 reads the composed frame back into page-locked host
    #      memory, every step; wall clock between barriers, max over ranks
    e2e_multi = None
    if world_size > 1:
        frame_dev = torch.empty((world_size * rows_per, W), dtype=torch.int32, device=dev) if rank == 0 else None
        host_frame = torch.empty((H, W), dtype=torch.int32).pin_memory() if rank == 0 else None
        gath = [frame_dev[r * rows_per:(r + 1) * rows_per] for r in range(world_size)] if rank == 0 else None

        def step_e2e():
            api.render_frame_into(batch, vp, cam.position, cfg_async, VD, my_stripe.data_ptr(), 0, ctx)
            with torch.cuda.stream(stream):
                dist.gather(my_stripe, gath, dst=0)
                if rank == 0:
                    host_frame.copy_(frame_dev[:H], non_blocking=True)
            ctx.synchronize()

        for _ in range(3):
            step_e2e()
        ne2e = max(20, min(K, 200))
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(ne2e):
            step_e2e()
        dist.barrier()
        torch.cuda.synchronize()
        el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
        e2e_multi = ne2e / float(el.item())

    if rank != 0:
        return finish_distributed(dist)  # waits for rank 0 (which still has the single-rank sections to run)

    # ---- warm-L2 back-to-back throughput (how the path is used in a render loop) --------------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nb2b = 200
    ctx.synchronize()
    e0.record(stream)
    for _ in range(nb2b):
        api.render_frame_device(batch, vp, cam.position, cfg_async, VD, ctx)  # this rank's rows only (no composite) when N > 1
    e1.record(stream)
    torch.cuda.synchronize()
    b2b_ms = e0.elapsed_time(e1) / nb2b

    # ---- per-kernel split (CUDA events inside the library), L2 flushed, 1 GPU worth of work ---------------------
    cfg_prof = api.VxFrameConfig.from_buffer_copy(cfg)
    cfg_prof.profile_kernels = 1
    ksum = np.zeros(4)
    nprof = 20
    for _ in range(nprof):
        flush_l2()
        api.render_frame_device(batch, vp, cam.position, cfg_prof, VD, ctx)
        ksum += api.frame_kernel_times(ctx)
    kms = ksum / nprof
    knames = ["frame_cull_kernel", "frame_setup_kernel", "(unused)", "frame_raster_kernel"]
    top = int(np.argmax(kms))
    st = api.frame_stats(ctx)

    # algorithmic bytes: framebuffer written once (colour u32 + depth f32, clear fused), visible quad
    # streams + mesh headers read once, chunk table read once + visibility written
    vis_ids_quads = st.n_quads
    rows_mine = cfg.stripe_rows if cfg.stripe_rows > 0 else H
    b_fb = W * rows_mine * 8
    b_quads = 3 * vis_ids_quads + 936 * st.n_survivors
    b_cull = 16 * n_chunks + 4 * n_chunks
    b_frame = b_fb + b_quads + b_cull
    peak, peak_src = measured_peaks()
    top_bytes = {0: b_cull, 1: b_quads, 2: 0, 3: b_fb + b_quads}[top]
    achieved = top_bytes / (kms[top] * 1e-3) / 1e9 if kms[top] > 0 else 0.0

    # ---- e2e through the host API: VP/camera in, ARGB frame out into pinned host memory, every step -------------
    # api.FrameLoop binds the framebuffer (device-mapped page-locked host memory the raster kernel writes in place) and the
    # draw list once; a frame is then one vx_render_frame call
    e2e_val = None
    h2d = 16 * 4 + 3 * 4 + C.sizeof(api.VxFrameConfig)
    d2h = W * H * 4 + 4 * n_chunks + 64  # frame + draw order + control block
    if world_size == 1:
        loop = api.FrameLoop(batch, cfg, view_distance=VD, want_depth=False, ctx=ctx)
        for _ in range(3):
            loop.render(vp, cam.position)
        ne2e = max(20, min(K, 200))
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.synchronize()
        t0 = time.perf_counter()
        s0.record(stream)
        for _ in range(ne2e):
            loop.render(vp, cam.position)
        s1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        e2e_ms = max(s0.elapsed_time(s1), wall * 1000.0) / ne2e
        e2e_val = 1000.0 / e2e_ms
        # extra (not the headline): the same frames double-buffered -- frame k is enqueued (vx_render_frame_into, async
        # submit, into one of two mapped host framebuffers) before the host waits for frame k-1, so launch latency and the
        # host wake-up hide behind the GPU; every frame still lands in host memory
        try:
            bufs = [loop.color, ctx.host_array((H, W), np.uint32)]
            bufs[1][...] = 0
            evs = [torch.cuda.Event(), torch.cuda.Event()]
            ptrs = [int(b.ctypes.data) for b in bufs]

            def submit(k):
                api.render_frame_into(batch, vp, cam.position, cfg_async, VD, ptrs[k & 1], 0, ctx)
                evs[k & 1].record(stream)

            for k in range(4):
                submit(k)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            submit(0)
            for k in range(1, ne2e):
                submit(k)
                evs[(k - 1) & 1].synchronize()
            evs[(ne2e - 1) & 1].synchronize()
            e2e_pipelined = ne2e / (time.perf_counter() - t0)
            if not np.array_equal(bufs[0], bufs[1]):
                e2e_pipelined = None
            api.frame_stats(ctx)  # surfaces any deferred scratch overflow of the async frames
        except Exception as ex:  # noqa: BLE001 -- an extra must never cost the headline line
            print("pipelined e2e skipped:", ex, file=sys.stderr)
            e2e_pipelined = None
    else:
        e2e_val = e2e_multi
        e2e_pipelined = None

    # ---- second BASELINE metric: chunks meshed / s (whole-world remesh sweep, inputs resident) -------------------
    def remesh():
        ctx.check(ctx.lib.vx_remesh_chunks_device(ctx.handle, C.c_void_p(d_vox.data_ptr()), C.c_void_p(d_nb.data_ptr()), None, batch.handle))

    for _ in range(3):
        remesh()
    ctx.synchronize()
    m_ms = []
    for _ in range(20):
        flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        remesh()
        b.record(stream)
        torch.cuda.synchronize()
        m_ms.append(a.elapsed_time(b))
    mesh_ms = float(np.mean(m_ms))
    chunks_per_s = n_chunks / (mesh_ms * 1e-3)
    n_nbr = int((nb >= 0).sum())
    b_mesh = n_chunks * (32768 + 792 + 144) + 1024 * n_nbr + 3 * total_quads
    mesh_gbs = b_mesh / (mesh_ms * 1e-3) / 1e9

    # large-batch meshing (BASELINE cfg 1 replicated: one terrain chunk, no neighbours, 16,384 copies = 512 MiB > L2), for the
    # chunk with the most quads of the world (worst case of the sweep) and for the median one
    rep = 16384
    qc_all = batch.download()["quad_count"]

    def big_batch(idx):
        d_big = d_vox[idx].repeat(rep, 1).contiguous()
        hb = C.c_void_p()
        ctx.check(ctx.lib.vx_mesh_chunks_device(ctx.handle, C.c_void_p(d_big.data_ptr()), None, None, None, rep, C.byref(hb)))
        big = api.MeshBatch(ctx, hb)
        quads = int(big.info().total_quads)
        bm = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            ctx.check(ctx.lib.vx_remesh_chunks_device(ctx.handle, C.c_void_p(d_big.data_ptr()), None, None, big.handle))
            b.record(stream)
            torch.cuda.synchronize()
            bm.append(a.elapsed_time(b))
        ms = float(np.mean(bm[1:]))
        big.release()
        del d_big
        return rep / (ms * 1e-3), (rep * (32768 + 792 + 144) + 3 * quads) / (ms * 1e-3) / 1e9, quads // rep

    big_cps, big_gbs, big_q = big_batch(int(np.argmax(qc_all)))
    med_cps, med_gbs, med_q = big_batch(int(np.argsort(qc_all)[len(qc_all) // 2]))

    # ---- meshing end to end in the steady state: the world's voxels come from page-locked host memory every sweep
    #      (one H2D copy into the resident device array), re-mesh into the existing batch, read the totals back
    mesh_e2e = None
    if world_size == 1:
        try:
            hv_t = torch.from_numpy(v).pin_memory()

            def sweep():
                with torch.cuda.stream(stream):
                    d_vox.copy_(hv_t, non_blocking=True)
                remesh()
                return batch.info().total_quads  # synchronises and reads the totals back

            for _ in range(3):
                sweep()
            t0 = time.perf_counter()
            reps = 20
            for _ in range(reps):
                tq_e2e = sweep()
            el = (time.perf_counter() - t0) / reps
            assert int(tq_e2e) == total_quads
            mesh_e2e = {"chunks_per_sec": n_chunks / el, "ms_per_world": el * 1e3, "h2d_bytes_per_sweep": int(v.nbytes), "d2h_bytes_per_sweep": 32,
                        "note": "818 x 32 KiB voxels H2D from page-locked memory + vx_remesh_chunks_device + totals read back, wall clock; PCIe-bound"}
        except Exception as e:
            mesh_e2e = {"error": repr(e)}

    # ---- terrain generated on the device and meshed without ever crossing PCIe: all 7,153 lattice chunks
    gen_mesh = None
    if world_size == 1:
        try:
            nb_full = world.neighbor_table()
            d_posf = torch.from_numpy(pos).to(dev)
            d_nbf = torch.from_numpy(nb_full).to(dev)
            d_voxf = torch.empty((pos.shape[0], 32768), dtype=torch.uint8, device=dev)
            tp = api.terrain_params()
            fl = api.generate_terrain(pos, d_voxf.data_ptr(), ctx, tp)
            d_flf = torch.from_numpy(fl).to(dev)
            hh = C.c_void_p()
            ctx.check(ctx.lib.vx_mesh_chunks_device(ctx.handle, C.c_void_p(d_voxf.data_ptr()), C.c_void_p(d_posf.data_ptr()),
                                                    C.c_void_p(d_nbf.data_ptr()), C.c_void_p(d_flf.data_ptr()), int(pos.shape[0]), C.byref(hh)))
            bfull = api.MeshBatch(ctx, hh)
            assert int(bfull.info().total_quads) == total_quads, "device-generated world meshes differently"
            ctx.synchronize()
            t0 = time.perf_counter()
            reps = 10
            for _ in range(reps):
                api.generate_terrain(pos, d_voxf.data_ptr(), ctx, tp)
                ctx.check(ctx.lib.vx_remesh_chunks_device(ctx.handle, C.c_void_p(d_voxf.data_ptr()), C.c_void_p(d_nbf.data_ptr()),
                                                          C.c_void_p(d_flf.data_ptr()), bfull.handle))
            ctx.synchronize()
            el = (time.perf_counter() - t0) / reps
            gen_mesh = {"chunks": int(pos.shape[0]), "varied_chunks": n_chunks, "ms_per_world": el * 1e3, "lattice_chunks_per_sec": pos.shape[0] / el,
                        "note": "vx_generate_terrain (7,153 positions -> voxels + Uniform flags on the device) + vx_remesh_chunks_device, wall clock"}
            bfull.release()
            del d_voxf
        except Exception as e:
            gen_mesh = {"error": repr(e)}

    # ---- adjacent rasterizers: the reference's own span-walker bench case and the macrotile frame ---
    a18 = None
    if world_size == 1:
        try:
            from oracle import binding as ob18
            sw_w, sw_h = 1920, 1080  # benches/span_walker.rs:36-77 "span_walker_full_packet_32_quads": a 4 x 8 grid of quads
            ii = np.arange(32)
            bx0 = (np.float32(-0.9) + (ii % 8).astype(np.float32) * np.float32(0.225)).astype(np.float32)
            by0 = (np.float32(-0.9) + (ii // 8).astype(np.float32) * np.float32(0.45)).astype(np.float32)
            bx1 = (bx0 + np.float32(0.2)).astype(np.float32)
            by1 = (by0 + np.float32(0.4)).astype(np.float32)
            bz = np.full(32, 0.5, dtype=np.float32)
            bt = ((ii % 4) + 1).astype(np.uint8)
            d_boxes = torch.from_numpy(np.concatenate([bx0, by0, bx1, by1, bz])).to(dev)
            d_types = torch.from_numpy(np.concatenate([bt, np.ones(32, dtype=np.uint8)])).to(dev)
            d_col = torch.zeros((sw_h, sw_w), dtype=torch.int32, device=dev)
            d_dep = torch.full((sw_h, sw_w), float("inf"), dtype=torch.float32, device=dev)

            def walk():
                ctx.check(ctx.lib.vx_span_walk_quads_device(ctx.handle, C.c_void_p(d_boxes.data_ptr()), C.c_void_p(d_types.data_ptr()), 32, sw_w, sw_h,
                                                            C.c_void_p(d_col.data_ptr()), C.c_void_p(d_dep.data_ptr())))

            walk()  # first call draws; the repeats below re-test equal depths like the reference's bench loop does
            ctx.synchronize()
            oc18 = np.zeros((sw_h, sw_w), dtype=np.uint32)
            od18 = np.full((sw_h, sw_w), np.inf, dtype=np.float32)
            ob18.span_walk_quads(oc18, od18, bx0, by0, bx1, by1, bz, bt)
            same18 = bool(np.array_equal(d_col.cpu().numpy().view(np.uint32), oc18))
            sw_ms = []
            for _ in range(10):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                walk()
                b.record(stream)
                torch.cuda.synchronize()
                sw_ms.append(a.elapsed_time(b))
            t0 = time.perf_counter()
            for _ in range(5):
                ob18.span_walk_quads(oc18, od18, bx0, by0, bx1, by1, bz, bt)
            sw_cpu_ms = (time.perf_counter() - t0) / 5 * 1e3
            # macrotile frame of the headline scene: caller's list = every meshed chunk that passes filter A
            vis18 = api.get_visible_chunks_frustum(p, cam.position, vp, VD, True, ctx)
            ids18 = np.flatnonzero((vis18 != 0) & (qc_all > 0)).astype(np.int32)
            cfg_m = api.VxFrameConfig.from_buffer_copy(cfg)
            cfg_m.macrotile = 1
            cfg_m.async_submit = 0
            mt_ms = []
            api.render_frame(batch, vp, cam.position, cfg_m, mesh_ids=ids18, ctx=ctx)
            for _ in range(10):
                t0 = time.perf_counter()
                api.render_frame(batch, vp, cam.position, cfg_m, mesh_ids=ids18, want_depth=False, color_out=loop.color, ctx=ctx)
                mt_ms.append((time.perf_counter() - t0) * 1e3)
            # Hyper-Pipeline (benches/differential_projection.rs:8-36 scene; host framebuffers in and out, wall clock)
            import vx_kat
            hb = api.BinaryGreedyMesher.mesh_batch(vx_kat.chunk_slab().reshape(1, -1), [(0, 0, 0)], None, None, ctx)
            from differential_projection_voxel_renderer_b200 import camera as _cam
            hvp = _cam.mat4_mul(_cam.perspective_rh(np.radians(np.float32(70.0)), 16 / 9, 0.1, 1000.0),
                                _cam.look_at_rh((64.0, 50.0, 100.0), (64.0, 32.0, 64.0), (0.0, 1.0, 0.0))).reshape(16)
            hfb = api.Framebuffer(1280, 720)
            hyper_quads = api.hyper_pipeline_render(hb, [0], hvp, hfb, ctx)
            hp_ms = []
            for _ in range(5):
                hfb = api.Framebuffer(1280, 720)
                t0 = time.perf_counter()
                api.hyper_pipeline_render(hb, [0], hvp, hfb, ctx)
                hp_ms.append((time.perf_counter() - t0) * 1e3)
            hb.release()
            a18 = {"hyper_pipeline_1280x720_one_chunk_ms": float(np.median(hp_ms)), "hyper_pipeline_visible_quads": hyper_quads,
                   "span_walker_full_packet_32_quads_1920x1080_ms": float(np.median(sw_ms)), "span_walker_cpu_port_ms": sw_cpu_ms,
                   "span_walker_matches_oracle": same18, "span_walker_launches": 5,
                   "macrotile_frame_1280x720_vd12_e2e_ms": float(np.median(mt_ms)), "macrotile_meshes": int(ids18.size),
                   "note": "vx_span_walk_quads_device (benches/span_walker.rs:36-77 workload, framebuffer resident) and vx_render_frame with cfg.macrotile = 1 (render_frame_macrotile) through the host API into mapped host memory, wall clock"}
        except Exception as e:  # noqa: BLE001
            a18 = {"error": repr(e)}

    # ---- BASELINE cfg 5 on this one GPU (context for the multi-GPU design point): 3840x2160, view distance 32 ------------
    cfg5 = None
    if world_size == 1:
        try:
            import vx_scenes
            pos5, world5, p5, v5, nb5 = vx_scenes.terrain_scene(32)
            batch5 = api.BinaryGreedyMesher.mesh_batch(v5, p5, nb5, None, ctx)
            cam5 = vx_scenes.main_camera(3840, 2160)
            vp5 = cam5.view_projection()
            c5 = api.default_frame_config(3840, 2160)
            api.render_frame_device(batch5, vp5, cam5.position, c5, 32, ctx)
            c5a = api.VxFrameConfig.from_buffer_copy(c5)
            c5a.async_submit = 1
            for _ in range(3):
                api.render_frame_device(batch5, vp5, cam5.position, c5a, 32, ctx)
            t5 = []
            for _ in range(20):
                flush_l2()
                a5, b5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a5.record(stream)
                api.render_frame_device(batch5, vp5, cam5.position, c5a, 32, ctx)
                b5.record(stream)
                torch.cuda.synchronize()
                t5.append(a5.elapsed_time(b5))
            st5 = api.frame_stats(ctx)
            ms5 = float(np.mean(t5))
            bytes5 = 3840 * 2160 * 8 + 3 * st5.n_quads + 936 * st5.n_survivors + 20 * int(p5.shape[0])
            cfg5 = {"workload": "3840x2160 view distance 32 (137,065 lattice chunks), 1 GPU, L2 flushed", "frame_ms": ms5, "frames_per_sec": 1000.0 / ms5,
                    "varied_chunks": int(p5.shape[0]), "visible_meshes": int(st5.n_survivors), "visible_quads": int(st5.n_quads),
                    "triangles": int(st5.n_triangles), "algorithmic_bytes": int(bytes5), "hbm_frac": bytes5 / (ms5 * 1e-3) / 1e9 / peak}
            batch5.release()
            del v5
        except Exception as e:  # never let the context line break the headline
            cfg5 = {"error": repr(e)}

    # ---- CPU baseline beside it (bounded sample) -------------------------------------------------------------------
    threads = os.cpu_count() or 1
    cpu_fps, cpu_n, cpu_el = cpu_frame_baseline(p, v, nb, cam, 10.0, threads)
    cpu_cps, cpu_mn, cpu_mel = cpu_mesh_baseline(v, nb, 3.0)

    out = {
        "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world_size, "steps": K, "warmup": max(3, args.warmup),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "full frame 1280x720 view distance 12 (filter A + filter B/sort + project/clip/cull + span raster), "
                               "meshes cached on device; BASELINE.json configs[2]",
                   "chunks": int(pos.shape[0]), "varied_chunks": n_chunks, "total_quads": total_quads,
                   "visible_meshes": int(st.n_survivors), "visible_quads": int(st.n_quads), "triangles": int(st.n_triangles),
                   "camera": "(0,10,20) yaw 0 pitch 0 fov 70", "projection": "exact (bit-identical to the CPU path)",
                   "l2": "flushed between timed frames (256 MiB device write, outside the timed events)",
                   "parallelism": "1 GPU" if world_size == 1 else f"{world_size} screen stripes of {rows_per} rows, NCCL gather to GPU0"},
        "clocks": clocks,
        "gpu_launches": int(l1 - l0),
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "note": ("per step every rank renders its stripe (vx_render_frame_device), NCCL gather to GPU0, rank 0 copies the composed ARGB frame to page-locked host memory; wall clock between barriers, max over ranks" if world_size > 1 else "api.FrameLoop.render -> vx_render_frame, one synchronous call per frame: VP + camera + config in; the ARGB frame lands in page-locked host memory (written over PCIe by the raster kernel itself, no staging copy) together with the draw order; the call returns after the stream has drained")},
        "roofline": {"bound": "hbm", "kernel": knames[top], "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic(knames[top]), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": int(top_bytes), "kernel_ms": float(kms[top]),
                     "kernel_share_of_step": float(kms[top] / kms.sum()) if kms.sum() > 0 else None},
        "cpu_baseline": {"value": cpu_fps, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{cpu_n} full 1280x720 vd12 frames in {cpu_el:.1f} s; C restatement of the reference CPU path "
                                   "(oracle/), stripe-parallel over all host threads (the Rust crate cannot be built here: no cargo)"},
        "extra": {
            "frame_ms_warm_l2_back_to_back": b2b_ms, "frames_per_sec_warm_l2": 1000.0 / b2b_ms,
            "kernel_ms": {k: float(x) for k, x in zip(knames, kms)},
            "launches_per_frame": int(launches_per_frame),
            "frame_algorithmic_bytes": int(b_frame), "frame_hbm_frac": (b_frame / (ms_per_step * 1e-3) / 1e9) / peak,
            "chunks_meshed_per_sec": chunks_per_s, "remesh_world_ms": mesh_ms, "remesh_world_chunks": n_chunks,
            "chunks_meshed_per_sec_sharded": sharded_chunks_per_s, "remesh_sharded_ms_max_over_ranks": shard_ms,
            "remesh_sharding": f"chunk id modulo {world_size} GPUs, voxels replicated, no collective",
            "remesh_algorithmic_GBps": mesh_gbs, "remesh_hbm_frac": mesh_gbs / peak,
            "chunks_meshed_per_sec_large_batch": big_cps, "large_batch": f"{rep} copies of the world's busiest terrain chunk ({big_q} quads), no neighbours (BASELINE configs[0] replicated, 512 MiB of voxels)",
            "chunks_meshed_per_sec_large_batch_median_chunk": med_cps, "median_chunk_quads": med_q,
            "large_batch_median_algorithmic_GBps": med_gbs, "large_batch_median_hbm_frac": med_gbs / peak,
            "large_batch_algorithmic_GBps": big_gbs, "large_batch_hbm_frac": big_gbs / peak,
            "cpu_chunks_meshed_per_sec_1_thread": cpu_cps,
            "cfg5_3840x2160_vd32": cfg5,
            "mesh_e2e_steady_state": mesh_e2e,
            "e2e_double_buffered_frames_per_s": round(e2e_pipelined, 1) if e2e_pipelined else None,
            "generate_and_mesh_on_device": gen_mesh,
            "adjacent_rasterizers_a18": a18,
            "frames_per_sec_alternate_frame_rendering": afr_fps,
            "alternate_frame_rendering": "N > 1 only: every GPU renders whole 1280x720 frames independently (no collective), total frames / max time over ranks; the headline value is the stripe-sharded single frame (strong scaling)",
            "reference_published": "162-168 fps on a 6-core i5-12400 (README.md:29-32)",
        },
    }
    emit_json_line(out)
    if dist is not None:
        return finish_distributed(dist)
    # The context (and with it the stream torch's allocators have recorded events on) is deliberately NOT torn down
    # here: freeing the torch tensors after the stream is gone aborts the process.  Leave at once; the driver owns
    # the process lifetime.
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


_REAL_STDOUT_FD = None


def emit_json_line(obj):
    """The ONE line of the contract, on the process's real stdout."""
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT_FD, line)


def main():
    # Libraries (NCCL's version banner, for one) print to file descriptor 1.  stdout must carry exactly one JSON line, so
    # fd 1 is pointed at stderr for the whole run and the result line is written to a saved copy of the real stdout.
    global _REAL_STDOUT_FD
    sys.stdout.flush()
    _REAL_STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_cuda(args)


if __name__ == "__main__":
    rc = main()
    # Leave without interpreter-exit destructors: torch frees cached device / page-locked tensors after the CUDA context
    # is already gone there and aborts the process (exit code 134) although the run has completed.
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(int(rc or 0))
