"""CPU, gloo, world_size 2: the host-side logic of the multi-GPU path (no kernels involved):
  * DDPPO._compute_var_mean's single packed all-reduce reproduces the reference's
    distributed_var_mean (habitat-baselines/habitat_baselines/rl/ddppo/algo/ddppo.py:59-84);
  * the flat-gradient exchange (SUM all-reduce + 1/world folded into the optimizer) equals DDP's
    unweighted gradient mean (test/test_ddppo_reduce.py:111-118: grads equal across ranks);
  * the packed RunningMeanAndVar statistics all-reduce equals the reference's three collectives."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import torch_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import habitat_lab_b200 as hb
        from habitat_lab_b200.rl.ppo import DDPPO

        torch.manual_seed(rank)
        adv = torch.randn(33, 5, 1) * (1 + rank) + 0.3 * rank
        fin = adv[torch.isfinite(adv)].double()
        stats = torch.stack([fin.sum(), (fin * fin).sum(), torch.tensor(float(fin.numel()), dtype=torch.float64),
                             torch.zeros((), dtype=torch.float64)])
        obj = DDPPO.__new__(DDPPO)
        obj._world, obj._group = world, None
        mean_var = DDPPO._fused_var_mean(obj, stats)
        var_r, mean_r = DDPPO._compute_var_mean(adv.flatten())   # the reference-signature hook gives the same numbers
        assert abs(float(mean_r) - float(mean_var[0])) < 1e-5 and abs(float(var_r) - float(mean_var[1])) < 1e-5
        # flat gradient exchange
        g = torch.randn(1000) + rank
        flat = g.clone()
        dist.all_reduce(flat)
        flat *= 1.0 / world
        # RunningMeanAndVar packed stats: per-channel (sum, sumsq) + frame count
        x = torch.rand(4, 4, 8, 8) + rank
        packed = torch.zeros(17, dtype=torch.float64)
        packed[:4] = x.double().sum((0, 2, 3))
        packed[8:12] = (x.double() ** 2).sum((0, 2, 3))
        packed[16] = x.shape[0]
        dist.all_reduce(packed)
        # DD-PPO preemption rule (ppo_trainer.py:641-653) over the shared store
        from habitat_lab_b200.rl.ppo_trainer import PPOTrainer, make_config

        tr = PPOTrainer(make_config(num_steps=128))
        tr._is_distributed = True
        store = dist.distributed_c10d._get_default_store()
        tr.num_rollouts_done_store = dist.PrefixStore("rollout_tracker", store)
        if rank == 0:
            tr.num_rollouts_done_store.set("num_done", "0")
        dist.barrier()
        early0 = (tr.should_end_early(31), tr.should_end_early(64))      # nobody done yet
        dist.barrier()
        tr.num_rollouts_done_store.add("num_done", 1)                     # both ranks finish their rollout
        dist.barrier()
        early1 = (tr.should_end_early(31), tr.should_end_early(32))      # 2 >= 0.6*2 but step must be >= 0.25*T
        packed = torch.cat([packed, torch.tensor([float(x) for x in early0 + early1], dtype=torch.float64)])
        q.put((rank, adv.numpy(), mean_var.numpy(), g.numpy(), flat.numpy(), x.numpy(), packed.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_host_logic():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda t: t[0])
    res = [(r[0],) + tuple(torch.from_numpy(a) for a in r[1:]) for r in res]
    [p.join(60) for p in procs]
    advs = [r[1] for r in res]
    var_ref, mean_ref = O.distributed_var_mean([a.flatten() for a in advs])
    for r in res:
        torch.testing.assert_close(r[2][0], mean_ref.float(), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(r[2][1], var_ref.float(), rtol=1e-5, atol=1e-6)
    mean_g = (res[0][3] + res[1][3]) / 2
    for r in res:
        torch.testing.assert_close(r[4], mean_g)          # same averaged gradient on every rank
    assert torch.equal(res[0][4], res[1][4])
    # equal per-rank batch sizes (the reference's assumption): packed stats == mean of means / vars
    xs = [r[5] for r in res]
    n_el = sum(x.shape[0] * 64 for x in xs)
    packed = res[0][6]
    for r in res:
        assert r[6][17:].tolist() == [0.0, 0.0, 0.0, 1.0], r[6][17:]
    mean = packed[:4] / n_el
    var = packed[8:12] / n_el - mean * mean
    cat = torch.cat(xs).transpose(1, 0).reshape(4, -1).double()
    torch.testing.assert_close(mean, cat.mean(-1), rtol=1e-9, atol=1e-12)
    torch.testing.assert_close(var, cat.var(-1, unbiased=False), rtol=1e-7, atol=1e-10)
