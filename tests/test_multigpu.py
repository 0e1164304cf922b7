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
        # completion without a collective: the launch is followed by a release store of sequence number 1 into this rank's word of
        # rank 0's buffer; rank 0 waits for every word on its own stream (bench.py does the same across NVLink)
        s.set_completion_flag(buf.flag_ptr(rank), 1)
        assert len(mine) > 0
        # round-robin shard in ONE launch: view v of this rank goes to global slot rank + v * world
        s.render_views_device(w, h, "longestaxis", [cams[i] for i in mine], buf.ptr_for(rank), view_stride=world)
        if rank == 0:
            status = torch.zeros(1, dtype=torch.int32, device="cuda:0")
            buf.wait_flags(torch.cuda.current_stream().cuda_stream, 1, timeout_ms=60000, d_status_ptr=status.data_ptr())
            torch.cuda.synchronize()
            assert int(status.item()) == 0, "timed out waiting for the completion words"
            gathered = buf.to_tensor().cpu().numpy()
            local = torch.zeros((n, h, w, 3), dtype=torch.uint8, device="cuda:0")
            s.render_views_device(w, h, "longestaxis", cams, local.data_ptr())
            s.synchronize()
            first = (bool(np.array_equal(gathered, local.cpu().numpy())), int(gathered.any(axis=(1, 2, 3)).sum()))
        s.set_completion_flag(None)
        # ---- the same batch with DYNAMICALLY claimed views (multigpu.render_views_dynamic: claim counter + frames-landed counter in rank 0's buffer)
        dist.barrier()
        if rank == 0:
            buf.reset_counters()
        dist.barrier()
        stream, claim_stream = torch.cuda.Stream(), torch.cuda.Stream()
        s.set_stream(stream.cuda_stream)
        mine_dyn = multigpu.render_views_dynamic(s, buf, list(reversed(cams)), w, h, "longestaxis", stream, claim_stream, scale=1)
        counts = [None, None]
        dist.all_gather_object(counts, mine_dyn)
        if rank == 0:
            status = torch.zeros(1, dtype=torch.int32, device="cuda:0")
            buf.wait_counter(stream.cuda_stream, n, timeout_ms=60000, d_status_ptr=status.data_ptr())
            torch.cuda.synchronize()
            assert int(status.item()) == 0, "timed out waiting for the frames-landed counter"
            gathered = buf.to_tensor().cpu().numpy()
            ok_dyn = bool(np.array_equal(gathered, local.flip(0).cpu().numpy())) and sum(counts) == n
            q.put((first[0] and ok_dyn, first[1]))
        s.reset_stream()
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


@pytest.mark.gpu
@pytest.mark.parametrize("storage,algo", [("vcs", "longestaxis"), ("hashtable", "original")])
def test_render_views_sharded_c_abi_one_process(storage, algo):
    """vrm_render_views_sharded (SURVEY.md 8b render_views(handles[], ...)): ONE process, several handles -- here two replicas on
    device 0, and as many more as the box has GPUs -- dynamic view claiming, frames gathered on the first handle's device.  Every
    frame equals the single render of its camera; every view is rendered exactly once; host and device outputs agree."""
    xyz, rgb = scenes.probe_scene()
    w, h, n = 320, 180, 9
    cams = [api.Camera((6.0 + 0.7 * i, 2.0 + 0.3 * i, 6.0 - 0.5 * i), (0.0, 0.0, -1.0), (0.0, 1.0, 0.0), 60.0, np.float32(w) / np.float32(h)) for i in range(n)]
    devices = [0, 0] + list(range(1, min(api.load_library().vrm_device_count(), 4)))
    replicas = []
    for d in devices:
        s = api.VoxelScene(d)
        s.add_voxels(xyz, rgb)
        s.generate_voxel_scene(storage)
        replicas.append(s)
    want = np.stack([replicas[0].render(w, h, algo, c, scale=8)["rgb"] for c in cams])
    got = api.render_views_sharded(replicas, w, h, algo, cams, scale=8)
    assert np.array_equal(got["rgb"], want)
    assert sum(got["views_per_scene"]) == n and len(got["views_per_scene"]) == len(replicas)
    dev_out = torch.zeros((n, h, w, 3), dtype=torch.uint8, device="cuda:0")
    got2 = api.render_views_sharded(replicas, w, h, algo, cams, scale=8, d_rgb_ptr=dev_out.data_ptr())
    assert got2["rgb"] is None and np.array_equal(dev_out.cpu().numpy(), want)
    one = api.render_views_sharded(replicas[:1], w, h, algo, cams[:2], scale=8)      # a single handle, fewer views than launch slots
    assert np.array_equal(one["rgb"], want[:2]) and one["views_per_scene"] == [2]
    pinned = torch.zeros((n, h, w, 3), dtype=torch.uint8).pin_memory()                # one handle + pinned frames: stored directly
    api.render_views_sharded(replicas[:1], w, h, algo, cams, scale=8, rgb_out=pinned.numpy())
    assert np.array_equal(pinned.numpy(), want)
    with pytest.raises(api.VrmError):
        api.render_views_sharded([replicas[0], replicas[0]], w, h, algo, cams, scale=8)
    for s in replicas:
        s.close()


@pytest.mark.gpu
def test_completion_flag_counts_render_launches():
    """vrm_scene_set_completion_flag: one sequence number per render launch, published behind the frame; vrm_wait_flags_device returns
    at once when the word is there and reports a time-out (status 1) when it is not."""
    xyz, rgb = scenes.probe_scene()
    s = api.VoxelScene(0)
    s.add_voxels(xyz, rgb)
    s.generate_voxel_scene("vcs")
    lib = api.load_library()
    w, h = 160, 90
    cam = api.Camera.reference_default(w, h)
    fb = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda:0")
    flag = torch.zeros(64, dtype=torch.int32, device="cuda:0")
    status = torch.zeros(1, dtype=torch.int32, device="cuda:0")
    s.set_stream(torch.cuda.current_stream().cuda_stream)
    s.set_completion_flag(flag.data_ptr(), 5)
    for algo in ("longestaxis", "original", "longestaxis"):
        s.render_device(w, h, algo, cam, fb.data_ptr(), scale=8)
    import ctypes as C
    assert lib.vrm_wait_flags_device(0, C.c_void_p(torch.cuda.current_stream().cuda_stream), C.c_void_p(flag.data_ptr()), 1, 64, 7, 5000, C.c_void_p(status.data_ptr())) == 0
    torch.cuda.synchronize()
    assert int(flag[0].item()) == 7 and int(status.item()) == 0
    assert lib.vrm_wait_flags_device(0, C.c_void_p(torch.cuda.current_stream().cuda_stream), C.c_void_p(flag.data_ptr()), 1, 64, 8, 50, C.c_void_p(status.data_ptr())) == 0
    torch.cuda.synchronize()
    assert int(status.item()) == 1                      # nobody publishes 8: the wait gives up after 50 ms
    s.set_completion_flag(None)
    s.render_device(w, h, "longestaxis", cam, fb.data_ptr(), scale=8)
    s.synchronize()
    assert int(flag[0].item()) == 7
    s.close()
