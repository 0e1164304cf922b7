"""Multi-GPU sharding of multi-view batches (BASELINE.json configs[3]).

Pixels and views are independent, so the path shards with NO data-path collective: the voxel structure is replicated
(every rank builds it on its own GPU from the same voxel list -- the build is deterministic), each rank renders a
contiguous block of the views in one launch (``vrm_render_views_device``), and the only exchange step is the gather of
the finished RGB8 frames on rank 0 (NCCL over NVLink on the GPU box; gloo in the CPU tests, where the renderer is a stub).

One process per GPU, ``torch.distributed`` for the plumbing.  The reference has no multi-GPU path at all
(main/Main.cu:82-94 pins device 0).
"""
from __future__ import annotations

from typing import Callable, Sequence


def shard_views(n_views: int, world_size: int, rank: int) -> range:
    """Contiguous block of view indices owned by ``rank``; the first ``n_views % world_size`` ranks get one more."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world size")
    base, extra = divmod(n_views, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def shard_views_interleaved(n_views: int, world_size: int, rank: int) -> range:
    """Views ``rank, rank + world_size, ...``: neighbouring views of an orbit cost about the same, so dealing them out round-robin
    balances the ranks better than contiguous blocks when the cost varies along the orbit.  With a shared frame buffer
    (``PeerFrameBuffer``) a rank renders its views in one launch with ``view_stride = world_size`` starting at slot ``rank``."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world size")
    return range(rank, n_views, world_size)


def gather_frames(local_frames, n_views: int, group=None, dst: int = 0):
    """Gather per-rank frame blocks ``[n_local, H, W, 3]`` (uint8, same device on every rank) on ``dst``.
    Returns the full ``[n_views, H, W, 3]`` tensor on ``dst`` and ``None`` elsewhere.  Ranks may own different numbers
    of views, so blocks are padded to the largest block for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = [len(shard_views(n_views, world, r)) for r in range(world)]
    biggest = max(counts)
    h, w = local_frames.shape[1:3]
    block = local_frames
    if block.shape[0] != biggest:
        pad = torch.zeros((biggest - block.shape[0], h, w, 3), dtype=torch.uint8, device=local_frames.device)
        block = torch.cat([block, pad], 0)
    block = block.contiguous()
    out = [torch.empty_like(block) for _ in range(world)] if rank == dst else None
    dist.gather(block, out, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([o[:c] for o, c in zip(out, counts)], 0)


def render_views_sharded(render_block: Callable[[Sequence, "object"], None], cameras: Sequence, width: int, height: int, device, group=None,
                         dst: int = 0):
    """Render ``cameras`` across the ranks of ``group`` and gather the frames on ``dst``.

    ``render_block(cams, out)`` must fill ``out[i]`` (``[len(cams), H, W, 3]`` uint8 on ``device``) with the frame of
    ``cams[i]`` -- on the GPU box it is ``lambda cams, out: scene.render_views_device(W, H, algo, cams, out.data_ptr())``.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = shard_views(len(cameras), world, rank)
    local = torch.zeros((len(mine), height, width, 3), dtype=torch.uint8, device=device)
    if len(mine):
        render_block([cameras[i] for i in mine], local)
    return gather_frames(local, len(cameras), group=group, dst=dst)


class PeerFrameBuffer:
    """The gathered frame buffer lives on rank ``dst`` and every other rank maps it through CUDA IPC, so that the render
    kernels store their frames straight into it over NVLink / NVSwitch (``vrm_peer_*`` in include/vrm_b200.h): the exchange
    is fused into the producing kernel instead of following it as a collective.  ``ptr_for(slot)`` is the device pointer a
    rank passes to ``VoxelScene.render_device`` / ``render_views_device`` for global frame slot ``slot``."""

    def __init__(self, n_frames: int, width: int, height: int, device: int, group=None, dst: int = 0):
        import ctypes as C

        import torch.distributed as dist
        from . import api

        self.lib = api.load_library()
        self.device, self.dst = device, dst
        self.rank = dist.get_rank(group)
        self.frame_bytes = width * height * 3
        self.shape = (n_frames, height, width, 3)
        self.owner = self.rank == dst
        self.ptr = C.c_void_p()
        # behind the frames: one completion word per rank, 256 bytes apart (vrm_scene_set_completion_flag / wait_flags)
        self.world = dist.get_world_size(group)
        self.flags_offset = (n_frames * self.frame_bytes + 255) // 256 * 256
        handle = (C.c_ubyte * 64)()
        if self.owner:
            # (+ two more words behind the per-rank flags: the view-claim counter and the frames-landed counter of render_views_dynamic)
            rc = self.lib.vrm_peer_alloc(device, self.flags_offset + 256 * (self.world + 2), C.byref(self.ptr), C.cast(handle, C.c_void_p))
            if rc:
                raise api.VrmError(f"vrm_peer_alloc failed: {rc}")
        box = [bytes(handle) if self.owner else None]
        dist.broadcast_object_list(box, src=dst, group=group)
        if not self.owner:
            raw = (C.c_ubyte * 64).from_buffer_copy(box[0])
            rc = self.lib.vrm_peer_open(device, C.cast(raw, C.c_void_p), C.byref(self.ptr))
            if rc:
                raise api.VrmError(f"vrm_peer_open failed: {rc} (no peer access between the GPUs?)")

    def ptr_for(self, slot: int) -> int:
        return self.ptr.value + slot * self.frame_bytes

    FLAG_STRIDE_WORDS = 64

    def flag_ptr(self, rank: int) -> int:
        """Device pointer of ``rank``'s completion word (zero after allocation)."""
        return self.ptr.value + self.flags_offset + 256 * rank

    def counter_ptr(self, which: int) -> int:
        """0: next unclaimed view (vrm_claim_next); 1: frames landed (vrm_scene_set_completion_counter).  Zero after allocation."""
        return self.ptr.value + self.flags_offset + 256 * (self.world + which)

    def reset_counters(self):
        """Owner only, between batches (synchronous)."""
        import torch
        zero = torch.zeros(128, dtype=torch.int32, device=f"cuda:{self.device}")
        import ctypes as C
        rc = self.lib.vrm_copy_device(self.device, C.c_void_p(self.counter_ptr(0)), C.c_void_p(zero.data_ptr()), 512)
        if rc:
            raise RuntimeError(f"vrm_copy_device failed: {rc}")

    def wait_counter(self, cuda_stream_ptr: int, n_frames: int, timeout_ms: int = 20000, d_status_ptr: int | None = None):
        """Owner only: enqueue a wait until ``n_frames`` frames have landed (counting completion signal)."""
        import ctypes as C
        rc = self.lib.vrm_wait_flags_device(self.device, C.c_void_p(cuda_stream_ptr), C.c_void_p(self.counter_ptr(1)), 1, 1, n_frames, timeout_ms,
                                            C.c_void_p(d_status_ptr or 0))
        if rc:
            raise RuntimeError(f"vrm_wait_flags_device failed: {rc}")

    def wait_flags(self, cuda_stream_ptr: int, min_value: int, timeout_ms: int = 20000, d_status_ptr: int | None = None):
        """Owner only: enqueue on ``cuda_stream_ptr`` a wait until every rank's completion word has reached ``min_value``."""
        import ctypes as C
        rc = self.lib.vrm_wait_flags_device(self.device, C.c_void_p(cuda_stream_ptr), C.c_void_p(self.flag_ptr(0)), self.world, self.FLAG_STRIDE_WORDS,
                                            min_value, timeout_ms, C.c_void_p(d_status_ptr or 0))
        if rc:
            raise RuntimeError(f"vrm_wait_flags_device failed: {rc}")

    def to_tensor(self):
        """Owner only: copy the gathered frames into a fresh torch tensor (after the producers have been synchronised)."""
        import ctypes as C

        import torch
        out = torch.empty(self.shape, dtype=torch.uint8, device=f"cuda:{self.device}")
        rc = self.lib.vrm_copy_device(self.device, C.c_void_p(out.data_ptr()), self.ptr, out.numel())
        if rc:
            raise RuntimeError(f"vrm_copy_device failed: {rc}")
        return out

    def close(self):
        if self.ptr and self.ptr.value:
            (self.lib.vrm_peer_free if self.owner else self.lib.vrm_peer_close)(self.device, self.ptr)
            self.ptr.value = None


def render_views_dynamic(scene, peer: PeerFrameBuffer, cameras: Sequence, width: int, height: int, algorithm, stream, claim_stream, scale=1,
                         translation=(0.0, 0.0, 0.0)) -> int:
    """One rank's part of a view batch whose views are claimed DYNAMICALLY from a counter in the gatherer's memory (BASELINE.json
    configs[3]: the views of an orbit differ in cost, static blocks or round-robin leave ranks idle at the end).  The rank keeps one
    frame rendering and one claimed: the claim for the next view runs on ``claim_stream`` while the current frame renders on
    ``stream``.  Every frame is stored by its kernel into slot ``view`` of ``peer`` and counted there (``peer.wait_counter`` on the
    owner).  Call ``peer.reset_counters()`` on the owner and a barrier before every batch.  Returns the number of views rendered here."""
    import ctypes as C
    import os

    import torch

    lib = peer.lib
    n = len(cameras)
    claimed = torch.zeros(2, dtype=torch.int32).pin_memory()
    scene.set_completion_counter(peer.counter_ptr(1))
    rendered = 0

    def claim(slot):
        rc = lib.vrm_claim_next(peer.device, C.c_void_p(claim_stream.cuda_stream), C.c_void_p(peer.counter_ptr(0)), C.c_void_p(claimed.data_ptr() + 4 * slot))
        if rc:
            raise RuntimeError(f"vrm_claim_next failed: {rc}")

    # A claim made while the previous frame still renders keeps the GPU busy, but the claimed view then waits behind that frame while
    # another rank may be idle: fine in the middle of a batch, costly at its end (64 views over 8 ranks: x6.7 instead of x7.4).  So the
    # claim overlaps the frame only while more than `tail` views are left; in the tail a rank claims when it is actually free (the claim
    # + launch gap, ~30 us, is idle time then -- against frames of milliseconds).
    tail = int(os.environ.get("VRM_CLAIM_TAIL", 2 * max(1, peer.world)))
    try:
        claim(0)
        k = 0
        while True:
            claim_stream.synchronize()
            v = int(claimed[k & 1].item())
            if v >= n:
                break
            scene.render_device(width, height, algorithm, cameras[v], peer.ptr_for(v), scale=scale, translation=translation)
            rendered += 1
            k += 1
            if v + tail >= n:
                stream.synchronize()
            claim(k & 1)          # fetched while the frame above renders, except in the tail
        stream.synchronize()
    finally:
        scene.set_completion_counter(None)
    return rendered
