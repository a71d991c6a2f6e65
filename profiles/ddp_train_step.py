"""Training config on > 1 GPU: DDP gradient all-reduce (NCCL over NVLink) around the drop-in modules.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \\
        profiles/ddp_train_step.py

Mirrors train.py:188-190 / model.py:84-86 of the reference: one process per GPU, DistributedDataParallel over
the aggregation network (train mode: BatchNorm batch statistics, autograd through the sm_100a forward/backward
kernels), Adam step.  Every rank gets a different shard of a fixed global batch.  Checks: (1) after backward the
gradients are identical on all ranks (the all-reduce happened), (2) they equal the average of the per-rank local
gradients computed without DDP, (3) parameters stay identical after the optimizer step, (4) the loss falls.
Prints one JSON line on rank 0 with the step time and the all-reduced bytes."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.nn.parallel import DistributedDataParallel as DDP  # noqa: E402

import aanet_b200.nets as n  # noqa: E402
from aanet_b200.sharding import shard_batch  # noqa: E402


def make():
    torch.manual_seed(326)
    agg = n.AdaptiveAggregation(64, num_deform_blocks=3, intermediate_supervision=True)
    for name, m in agg.named_modules():
        if name.endswith("offset_conv"):
            torch.nn.init.normal_(m.weight, std=0.05)
    return agg


def loss_fn(agg, L, R, target):
    outs = agg(n.CostVolumePyramid(64)(L, R))
    est = n.DisparityEstimation(64)
    return sum(torch.nn.functional.smooth_l1_loss(est(o), target[s]) for s, o in enumerate(outs))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(7)
    GB, C, H, W = 2 * world, 32, 48, 96                         # global batch, sharded contiguously
    L = [torch.relu(torch.randn(GB, C, H >> s, W >> s, generator=g)) for s in range(3)]
    R = [torch.relu(torch.randn(GB, C, H >> s, W >> s, generator=g)) for s in range(3)]
    T = [torch.rand(GB, H >> s, W >> s, generator=g) * (63 >> s) for s in range(3)]
    Ll, Rl, Tl = [[t.to(dev) for t in shard_batch(x, rank, world)] for x in (L, R, T)]

    # local gradients without DDP (reference for the all-reduce)
    ref = make().to(dev).train()
    loss_fn(ref, Ll, Rl, Tl).backward()
    local = torch.cat([p.grad.flatten() for p in ref.parameters() if p.grad is not None])
    mean = local.clone()
    dist.all_reduce(mean)
    mean /= world

    net = DDP(make().to(dev).train(), device_ids=[dev.index])
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    losses, times = [], []
    for step in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = loss_fn(net, Ll, Rl, Tl)
        loss.backward()
        if step == 0:
            got = torch.cat([p.grad.flatten() for p in net.module.parameters() if p.grad is not None])
            err = float((got - mean).abs().max() / mean.abs().max())
            gathered = [torch.empty_like(got) for _ in range(world)]
            dist.all_gather(gathered, got)
            same = all(torch.equal(gathered[0], x) for x in gathered)
        opt.step()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        lt = loss.detach().clone()
        dist.all_reduce(lt)
        losses.append(float(lt) / world)
    flat = torch.cat([p.detach().flatten() for p in net.module.parameters()])
    allp = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(allp, flat)
    params_same = all(torch.equal(allp[0], x) for x in allp)
    if rank == 0:
        print(json.dumps({"world": world, "global_batch": GB, "grad_allreduce_bytes": int(got.numel() * 4),
                          "grads_identical_across_ranks": same, "ddp_vs_mean_of_local_grads_rel": err,
                          "params_identical_after_steps": params_same, "loss": losses,
                          "ms_per_step_last3": 1e3 * sum(times[-3:]) / 3}))
    ok = same and params_same and err < 1e-4 and losses[-1] < losses[0]
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
