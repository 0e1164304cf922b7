"""BASELINE.json configs[3] across GPUs: the 64-view 1920x1080 orbit of the 2048^3 sparse-shell scene, views sharded in
contiguous blocks over the ranks (one process per GPU, torchrun), each rank renders its block in ONE launch straight into
rank 0's frame buffer over NVLink peer memory (multigpu.PeerFrameBuffer); one 4-byte all-reduce tells rank 0 the frames are
complete.  Prints one JSON line on rank 0: ms for the whole orbit (max over ranks, CUDA events), Mrays/s, per-frame ms.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/orbit_multi.py [--algo longestaxis]
"""
import argparse, json, os, sys
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxelraymarcher_b200 import api, multigpu


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--algo", default="longestaxis")
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--interleave", action="store_true", help="deal the views out round-robin (view_stride = world size) instead of contiguous blocks")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    w, h, views, size = 1920, 1080, 64, a.size
    s = api.VoxelScene(local)
    n = s.generate_sparse_shells(size, 64, 7, 35)          # replicated: every rank generates and builds the same scene on its own GPU
    build_ms = s.generate_voxel_scene("vcs")
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    s.set_stream(stream.cuda_stream)
    cams = []
    for v in range(views):
        ang = 2.0 * np.pi * (v + 0.37) / views
        r, el, c = 1.5 * size / 2, np.deg2rad(20.0), size / 2
        cams.append(api.Camera((float(c + r * np.cos(el) * np.cos(ang)), float(c + r * np.sin(el)), float(c + r * np.cos(el) * np.sin(ang))), (c, c, c), (0.0, 1.0, 0.0), 60.0, np.float32(w) / np.float32(h)))
    mine = multigpu.shard_views_interleaved(views, world, rank) if a.interleave else multigpu.shard_views(views, world, rank)
    stride = world if a.interleave else 1
    peer = multigpu.PeerFrameBuffer(views, w, h, local) if world > 1 else None
    local_frames = torch.zeros((len(mine), h, w, 3), dtype=torch.uint8, device=dev) if peer is None else None
    out_ptr = peer.ptr_for(mine.start) if peer is not None else local_frames.data_ptr()
    token = torch.zeros(1, dtype=torch.int32, device=dev)
    times = []
    for it in range(a.iters + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        s.render_views_device(w, h, a.algo, [cams[i] for i in mine], out_ptr, view_stride=stride if peer is not None else 1)
        if world > 1:
            dist.all_reduce(token)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it > 0:
            times.append(float(t.item()))
    ok = None
    if peer is not None:
        # the gathered block of this rank must equal a local render of the same views
        check = torch.zeros((len(mine), h, w, 3), dtype=torch.uint8, device=dev)
        s.render_views_device(w, h, a.algo, [cams[i] for i in mine], check.data_ptr())
        torch.cuda.synchronize(dev)
        dist.barrier()
        if rank == 0:
            full = peer.to_tensor()
            ok = bool(torch.equal(full[mine.start:mine.stop:mine.step], check))
    if rank == 0:
        ms = float(np.median(times))
        print(json.dumps(dict(config="4: 2048^3 sparse shells, 64-view orbit 1920x1080", algo=a.algo, n_gpus=world, views=views, voxels=n, build_ms=build_ms,
                              sharding="interleaved" if a.interleave else "contiguous", ms_orbit=ms, ms_per_frame=ms / views, mrays_per_s=w * h * views / ms / 1e3, rank0_block_verified=ok)), flush=True)
    if world > 1:
        dist.barrier()
        if peer is not None:
            peer.close()
        dist.destroy_process_group()
    s.close()


if __name__ == "__main__":
    main()
