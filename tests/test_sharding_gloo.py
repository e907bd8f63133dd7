"""world_size-2 gloo test (CPU) of the N>1 path: LPT sharding + result gather.  The compute
itself needs a GPU, so each rank produces deterministic stand-in picks; what is tested is that
every video is processed exactly once and the gathered summaries are identical on all ranks and
equal to the single-process result."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import avsum_b200  # noqa: F401
from avsum_b200 import sharding, synth


def _fake_picks(i, S):
    return (np.random.default_rng(i).random(S) < 0.3).astype(np.uint8)


def _worker(rank, world, port, lengths, n_segs, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shards = sharding.shard_videos(lengths, world)
    mine = shards[rank]
    picks = [_fake_picks(i, n_segs[i]) for i in mine]
    full = sharding.gather_picks(mine, picks, len(lengths))
    ret[rank] = [p.tolist() for p in full]
    dist.barrier()
    dist.destroy_process_group()


def test_shard_videos_is_a_balanced_partition():
    lengths = [v.T for v in synth.config2()]
    for world in (1, 2, 4, 8):
        shards = sharding.shard_videos(lengths, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(len(lengths)))
        loads = [sum(sharding.video_cost(lengths[i]) for i in s) for s in shards]
        assert max(loads) <= 1.15 * (sum(loads) / world)
    assert sharding.shard_videos([], 4) == [[], [], [], []]
    assert sharding.shard_videos([5], 2) == [[0], []]


def test_two_rank_gather_equals_single_process():
    vids = synth.config2()[:11]
    lengths = [v.T for v in vids]
    n_segs = [int(v.cps.shape[0]) for v in vids]
    want = [_fake_picks(i, n_segs[i]).tolist() for i in range(len(vids))]
    single = sharding.gather_picks(list(range(len(vids))), [np.asarray(w, np.uint8) for w in want], len(vids))
    assert [p.tolist() for p in single] == want
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, lengths, n_segs, ret), nprocs=2, join=True)
    assert ret[0] == want and ret[1] == want


def _grad_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from avsum_b200.training import allreduce_gradients
    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
    for i, p in enumerate(lin.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    n = allreduce_gradients(lin.parameters())
    ret[rank] = (n, [float(p.grad.flatten()[0]) for p in lin.parameters()])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_averages_one_flat_bucket():
    """Data-parallel training (BASELINE configs[4]): one flat bucket, sum over ranks, divided by world."""
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_grad_worker, args=(2, port, ret), nprocs=2, join=True)
    want = [1.5 * (i + 1) for i in range(4)]          # mean of (1, 2) * (i + 1)
    assert ret[0][0] == ret[1][0] == 5 * 3 + 3 + 3 * 2 + 2
    assert ret[0][1] == want and ret[1][1] == want


def _bucket_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from avsum_b200.training import GradBuckets
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 4), torch.nn.Tanh(),
                              torch.nn.Linear(4, 1))
    unused = torch.nn.Parameter(torch.ones(3))             # never reaches the loss: must end with a zero gradient
    params = list(net.parameters()) + [unused]
    gb = GradBuckets(params, bucket_mb=30 * 4 / (1 << 20))    # ~30 elements per bucket -> several buckets
    x = torch.randn(7, 6, generator=torch.Generator().manual_seed(10 + rank))
    out = []
    for _ in range(2):                                      # two steps: the hooks re-arm
        for p in params:
            p.grad = None
        gb.start()
        net(x).pow(2).mean().backward()
        n = gb.finish()
        out.append([p.grad.clone() for p in params])
    ret[rank] = (n, len(gb.buckets), [[g.tolist() for g in step] for step in out],
                 all(p.grad.data_ptr() == gb.slot[id(p)][0].data_ptr() for p in params))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_bucketed_overlapped_allreduce_equals_the_mean_of_local_gradients():
    """training.GradBuckets: gradients land in one flat buffer (reverse parameter order), buckets are all-reduced
    asynchronously from the autograd hooks, finish() averages -- equal to the mean of the two ranks' local gradients,
    on both ranks, for two consecutive steps; a parameter without gradient gets zeros."""
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_bucket_worker, args=(2, port, ret), nprocs=2, join=True)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 4), torch.nn.Tanh(),
                              torch.nn.Linear(4, 1))
    local = []
    for rank in range(2):
        net.zero_grad()
        x = torch.randn(7, 6, generator=torch.Generator().manual_seed(10 + rank))
        net(x).pow(2).mean().backward()
        local.append([p.grad.clone() for p in net.parameters()])
    want = [(a + b) / 2 for a, b in zip(*local)] + [torch.zeros(3)]
    assert ret[0][0] == ret[1][0] == sum(w.numel() for w in want)
    assert ret[0][1] >= 3 and ret[0][3] and ret[1][3]       # several buckets; p.grad are views of the flat buffer
    for rank in range(2):
        for step in ret[rank][2]:
            for g, w in zip(step, want):
                assert torch.allclose(torch.tensor(g), w, rtol=1e-6, atol=1e-7)
