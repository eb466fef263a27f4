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


def _pair_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from types import SimpleNamespace
    from ctrlv_b200 import parallel
    r, w, _ = parallel.init_from_env(backend="gloo")
    pair = parallel.CfgPair(r, w)
    # stand-in for the model output of this rank's CFG branch of clip `pair.pair`
    local = torch.full((6, 4), float(10 * pair.pair + pair.branch)) + torch.arange(4.0)
    step = SimpleNamespace(noise=torch.zeros(12, 4), noise_local=local)
    pair.exchange(step)
    torch.save(dict(pair=pair.pair, branch=pair.branch, noise=step.noise), os.path.join(out_dir, f"p{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_cfg_pair_exchange_world4(tmp_path):
    """CFG-branch sharding: ranks (2k, 2k+1) form pair k; the exchange leaves both with
    [uncond half | cond half] of their own clip only."""
    world = 4
    mp.spawn(_pair_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        got = torch.load(os.path.join(tmp_path, f"p{r}.pt"))
        k = r // 2
        assert (got["pair"], got["branch"]) == (k, r % 2)
        want = torch.cat([torch.full((6, 4), float(10 * k + b)) + torch.arange(4.0) for b in (0, 1)])
        assert torch.equal(got["noise"], want)


def test_time_context_rotation_matches_s_major_order():
    """The rotation a branch-sharded process applies to the whole-batch context table reproduces the
    diffusers-0.27.2 pairing row (b, s) <-> context (b*S + s) % 2B."""
    from ctrlv_b200 import parallel
    for B in (1, 2, 3):
        for S in (1, 3, 40, 45, 2560):
            table = torch.arange(2 * B)  # context ids of the whole CFG batch
            for branch in (0, 1):
                rot = parallel.time_context_rotation(branch, B, S)
                local = torch.roll(table, shifts=-rot, dims=0)
                for bl in range(B):
                    for s in (0, 1, 2, S - 1):
                        want = ((branch * B + bl) * S + s) % (2 * B)
                        assert int(local[(bl * S + s) % (2 * B)]) == want
    assert parallel.cfg_pair_of(5) == (2, 1)
