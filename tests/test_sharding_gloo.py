"""N > 1 host logic on the CPU: world_size 2 over gloo (127.0.0.1)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rlao_b200 import sharding


def test_shard_envs_partitions_without_overlap():
    for total, ws in [(8192, 8), (1024, 1), (10, 4), (7, 8)]:
        seen = []
        for r in range(ws):
            n, off = sharding.shard_envs(total, r, ws)
            seen += list(range(off, off + n))
        assert seen == list(range(total))


def _worker(rank, ws, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        n, off = sharding.shard_envs(10)
        sr, rew, cnt = sharding.reduce_rollout_stats(strehl_sum=0.5 * n, reward_sum=-2.0 * n, count=n)
        states = torch.full((3, 2, 2), float(rank))
        actions = torch.full((3, 1), float(rank) + 10)
        gs, ga = sharding.gather_windows(states, actions)
        # the env's episode average reduces over ranks too (rank r holds Strehl values r+1)
        from rlao_b200.OOPAOEnv.OOPAOEnvRazor import OOPAO
        env = OOPAO()
        env.SR = [torch.full((4,), float(rank + 1)), torch.full((4,), float(rank + 1))]
        avg = env.calculate_strehl_AVG()
        # PO4AO trainers average gradients over ranks: each rank trains on its own (different) replay shard, both
        # end up with identical weights
        from rlao_b200.PO4AO import mbrl
        from rlao_b200.PO4AO.conv_models_simple import EnsembleDynamics
        from rlao_b200.PO4AO.util_simple import EfficientExperienceReplay
        torch.manual_seed(0)
        xv, yv = torch.nonzero(torch.ones(5, 5), as_tuple=True)
        dyn = EnsembleDynamics(xv, yv, 2, n_models=1)
        rp = EfficientExperienceReplay((5, 5), (5, 5), max_size=64)
        gen = torch.Generator().manual_seed(100 + rank)
        for _ in range(16):
            rp.append(torch.randn(5, 5, generator=gen), torch.randn(5, 5, generator=gen), 0.0, torch.randn(5, 5, generator=gen))
        mbrl.train_dynamics(2, 8, 4, dyn, torch.optim.SGD(dyn.parameters(), lr=0.1), rp, dyn_iters=2, device="cpu")
        wsum = float(sum(p.double().abs().sum() for p in dyn.parameters()))
        q.put((rank, n, off, sr, rew, cnt, gs[:, 0, 0].tolist(), ga[:, 0].tolist(), avg, wsum))
    finally:
        dist.destroy_process_group()


def test_world_size_two_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, n0, o0, sr0, rew0, c0, gs0, ga0, avg0, w0), (r1, n1, o1, sr1, rew1, c1, gs1, ga1, avg1, w1) = res
    assert (n0, o0, n1, o1) == (5, 0, 5, 5)
    assert sr0 == sr1 == 0.5 and rew0 == rew1 == -2.0 and c0 == c1 == 10
    assert gs0 == gs1 == [0.0, 0.0, 0.0, 1.0, 1.0, 1.0]
    assert ga0 == ga1 == [10.0, 10.0, 10.0, 11.0, 11.0, 11.0]
    assert avg0 == avg1 == 1.5
    assert w0 == w1
