"""world_size-2 gloo test (CPU) of the sample-parallel host logic: shard, run, all-gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_clips, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from ctrlv_b200 import parallel
    r, w, _ = parallel.init_from_env(backend="gloo")
    mine = parallel.shard_clips(n_clips, w, r)
    # stand-in for "sample clip i": a deterministic function of the clip index only
    local = torch.stack([torch.full((2, 4, 3, 5), float(i)) + torch.arange(5.0) for i in mine]) if mine else torch.zeros(0, 2, 4, 3, 5)
    allv = parallel.gather_latents(local, n_clips, w, r)
    torch.save(allv, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_matches_single_process(tmp_path):
    n_clips, world = 5, 2
    mp.spawn(_worker, args=(world, _free_port(), n_clips, str(tmp_path)), nprocs=world, join=True)
    ref = torch.stack([torch.full((2, 4, 3, 5), float(i)) + torch.arange(5.0) for i in range(n_clips)])
    for r in range(world):
        got = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert torch.equal(got, ref)
