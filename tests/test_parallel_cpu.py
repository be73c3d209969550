"""CPU, world_size 2 over gloo: the host-side data-parallel logic (parallel.py) -- flat-bucket gradient
all-reduce equals the single-process full-batch gradient, shard ranges tile the batch, and the global
q-regulariser denominator makes the rank-averaged loss equal the full-batch loss (checked with the
numpy oracle of the discriminative loss)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from isa_b200 import parallel, synth
    from oracle import disc_loss as O
    r, w, _ = parallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, world) and parallel.is_distributed()
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    for p in net.parameters():
        dist.broadcast(p.data, src=0)
    X = torch.randn(8, 6)
    Y = torch.randn(8, 3)
    lo, hi = parallel.shard_range(8, rank, world)
    bucket = parallel.FlatGradBucket(net.parameters())
    for step in range(2):           # second step exercises "grads alias the bucket"
        bucket.zero()
        loss = ((net(X[lo:hi]) - Y[lo:hi]) ** 2).mean()
        loss.backward()
        bucket.allreduce(average=True)
    flat = bucket.flat.clone()
    # q-regulariser denominator: rank-averaged local losses == full-batch oracle loss
    d = synth.batch(3, 4, 8, 16, 16, 4, n_min=1, n_max=4, fg_frac=0.2 + 0.5 * 0)
    d["labels"][0][:] = 255
    d["labels"][0][:3, :3] = 0      # very different foreground counts per image
    d["n_objects"][0] = 1
    l0, h0 = parallel.shard_range(4, rank, world)
    local_fg = torch.tensor(float((d["labels"][l0:h0] != 255).sum()))
    qden = parallel.global_q_denominator(local_fg)
    o = O.discriminative_loss(d["emb"][l0:h0], d["labels"][l0:h0], d["n_objects"][l0:h0], 4, 0.5, 1.5, 2)
    var_t, _, _, qreg_local = o["terms"]
    q_sum = qreg_local * float((d["labels"][l0:h0] != 255).sum())
    local_loss = torch.tensor([var_t + 0.005 * q_sum / float(qden)], dtype=torch.float64)
    dist.all_reduce(local_loss)
    local_loss /= world
    if rank == 0:
        torch.save({"flat": flat, "loss": local_loss, "net": net.state_dict()}, os.path.join(out_dir, "r0.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_bucket_allreduce_and_global_q_denominator(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = torch.load(os.path.join(str(tmp_path), "r0.pt"))
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    net.load_state_dict(got["net"])
    X = torch.randn(8, 6)
    Y = torch.randn(8, 3)
    ((net(X) - Y) ** 2).mean().backward()
    want = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    assert torch.allclose(got["flat"], want, atol=1e-6)
    from isa_b200 import synth
    from oracle import disc_loss as O
    d = synth.batch(3, 4, 8, 16, 16, 4, n_min=1, n_max=4, fg_frac=0.2)
    d["labels"][0][:] = 255
    d["labels"][0][:3, :3] = 0
    d["n_objects"][0] = 1
    full = O.discriminative_loss(d["emb"], d["labels"], d["n_objects"], 4, 0.5, 1.5, 2)
    assert abs(float(got["loss"]) - float(full["loss"])) < 1e-9 * max(1.0, abs(float(full["loss"])))


def _lr_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from isa_b200 import parallel
    parallel.init_from_env(backend="gloo")
    p = torch.nn.Parameter(torch.zeros(3))
    opt = torch.optim.SGD([p], lr=1.0)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode='min', factor=0.5, patience=1)
    local_sched_opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(3))], lr=1.0)
    local_sched = torch.optim.lr_scheduler.ReduceLROnPlateau(local_sched_opt, mode='min', factor=0.5, patience=1)
    lrs, local_lrs, costs = [], [], []
    for epoch in range(6):
        # rank 0's shard keeps improving, rank 1's plateaus: rank-local costs would drop the lr on rank 1 only
        local = {"Cost": (1.0 / (epoch + 1) if rank == 0 else 1.0 + 0.01 * epoch) * 3, "INS Cost": float(rank)}
        m = parallel.allreduce_epoch_metrics(local, 3 if rank == 0 else 5, torch.device("cpu"))
        sched.step(m["Cost"])
        local_sched.step(local["Cost"] / 3.0)
        lrs.append(opt.param_groups[0]["lr"])
        local_lrs.append(local_sched_opt.param_groups[0]["lr"])
        costs.append(m["Cost"])
    torch.save({"lrs": lrs, "local_lrs": local_lrs, "costs": costs}, os.path.join(out_dir, "lr%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_epoch_metrics_allreduce_keeps_the_learning_rate_identical_across_ranks(tmp_path):
    """ADVICE (round 1): ReduceLROnPlateau must see the same validation cost on every rank."""
    world = 2
    mp.spawn(_lr_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a = torch.load(os.path.join(str(tmp_path), "lr0.pt"))
    b = torch.load(os.path.join(str(tmp_path), "lr1.pt"))
    assert a["costs"] == b["costs"] and a["lrs"] == b["lrs"]
    # sample-weighted mean over both shards: (3 * c0 + 5 * c1) / 8 with the sums as handed over
    assert abs(a["costs"][0] - (3.0 + 3.0) / 8.0) < 1e-12
    assert a["local_lrs"] != b["local_lrs"], "the scenario must be one where rank-local costs WOULD diverge"


def test_shard_range_tiles_the_batch():
    from isa_b200 import parallel
    for gb in (1, 7, 16, 128):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(gb, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
