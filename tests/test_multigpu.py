"""N>1 path: view sharding + frame gather.  CPU tier: world_size-2 gloo with a stub renderer (the collective plumbing and
the shard arithmetic); GPU tier: two real scenes on one GPU emulate two ranks' blocks and are compared with a single
launch over all views, and two PROCESSES on one GPU run the fused peer-memory exchange through real CUDA IPC (a 2-rank NCCL run
needs two GPUs: bench.py --gpus 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.common import scenes
from voxelraymarcher_b200 import api, multigpu


def test_shard_views_interleaved_partitions_exactly():
    for n in (0, 1, 5, 64, 67):
        for world in (1, 2, 3, 8):
            seen = sorted(v for r in range(world) for v in multigpu.shard_views_interleaved(n, world, r))
            assert seen == list(range(n))
            sizes = [len(multigpu.shard_views_interleaved(n, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        multigpu.shard_views_interleaved(4, 2, 2)


def test_shard_views_partitions_exactly():
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            blocks = [multigpu.shard_views(n, world, r) for r in range(world)]
            flat = [i for b in blocks for i in b]
            assert flat == list(range(n))
            assert max(len(b) for b in blocks) - min(len(b) for b in blocks) <= 1
    with pytest.raises(ValueError):
        multigpu.shard_views(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_views, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    h, w = 6, 8
    cams = list(range(n_views))  # the stub renderer only needs something to identify the view

    def render_block(block, out):
        for i, view in enumerate(block):
            out[i] = torch.full((h, w, 3), view + 1, dtype=torch.uint8)
            out[i, 0, 0, 0] = rank  # who rendered it

    full = multigpu.render_views_sharded(render_block, cams, w, h, torch.device("cpu"))
    if rank == 0:
        q.put(full.numpy())
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_views", [5, 8])
def test_render_views_sharded_gloo_world2(n_views):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_views, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert full.shape == (n_views, 6, 8, 3)
    owners = [0 if v in multigpu.shard_views(n_views, 2, 0) else 1 for v in range(n_views)]
    for v in range(n_views):
        assert full[v, 1, 1, 0] == v + 1          # the right view in the right slot
        assert full[v, 0, 0, 0] == owners[v]       # rendered by the rank that owns it


@pytest.mark.gpu
def test_sharded_blocks_equal_single_launch():
    xyz, rgb = scenes.probe_scene()
    w, h = 320, 180
    cams = [api.Camera((6.0 + 0.7 * i, 2.0 + 0.3 * i, 6.0 - 0.5 * i), (0.0, 0.0, -1.0), (0.0, 1.0, 0.0), 60.0, np.float32(w) / np.float32(h)) for i in range(5)]
    replicas = []
    for _ in range(2):   # "ranks": the structure is replicated, each builds its own copy
        s = api.VoxelScene(0)
        s.add_voxels(xyz, rgb)
        s.generate_voxel_scene("vcs")
        replicas.append(s)
    whole = torch.zeros((5, h, w, 3), dtype=torch.uint8, device="cuda:0")
    replicas[0].render_views_device(w, h, "longestaxis", cams, whole.data_ptr())
    replicas[0].synchronize()
    parts = []
    for rank, s in enumerate(replicas):
        mine = multigpu.shard_views(5, 2, rank)
        out = torch.zeros((len(mine), h, w, 3), dtype=torch.uint8, device="cuda:0")
        s.render_views_device(w, h, "longestaxis", [cams[i] for i in mine], out.data_ptr())
        s.synchronize()
        parts.append(out)
    assert torch.equal(torch.cat(parts, 0), whole)
    assert whole.any()


def _peer_worker(rank, world, port, q):
    """Two processes on ONE GPU: rank 1 maps rank 0's frame buffer through CUDA IPC and its render kernel stores into it."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        xyz, rgb = scenes.probe_scene()
        w, h, n = 320, 180, 5
        cams = [api.Camera((6.0 + 0.7 * i, 2.0 + 0.3 * i, 6.0 - 0.5 * i), (0.0, 0.0, -1.0), (0.0, 1.0, 0.0), 60.0, np.float32(w) / np.float32(h)) for i in range(n)]
        s = api.VoxelScene(0)
        s.add_voxels(xyz, rgb)
        s.generate_voxel_scene("vcs")          # the structure is replicated: every rank builds its own
        buf = multigpu.PeerFrameBuffer(n, w, h, 0)
        mine = multigpu.shard_views_interleaved(n, world, rank)
        if mine:
            # round-robin shard in ONE launch: view v of this rank goes to global slot rank + v * world
            s.render_views_device(w, h, "longestaxis", [cams[i] for i in mine], buf.ptr_for(rank), view_stride=world)
        s.synchronize()
        dist.barrier()                          # (bench.py uses a 4-byte all-reduce for the same purpose)
        if rank == 0:
            gathered = buf.to_tensor().cpu().numpy()
            local = torch.zeros((n, h, w, 3), dtype=torch.uint8, device="cuda:0")
            s.render_views_device(w, h, "longestaxis", cams, local.data_ptr())
            s.synchronize()
            q.put((bool(np.array_equal(gathered, local.cpu().numpy())), int(gathered.any(axis=(1, 2, 3)).sum())))
        dist.barrier()                          # the owner frees the buffer only after every rank is done with it
        buf.close()
        s.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_fused_peer_memory_exchange_two_processes_one_gpu():
    """vrm_peer_alloc / vrm_peer_open / vrm_copy_device / vrm_peer_close / vrm_peer_free with real CUDA IPC between two processes
    (both on GPU 0, so it runs on a one-GPU box; bench.py --gpus N is the same code across NVLink)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    equal, frames = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert equal and frames == 5
