"""Where the data-parallel step spends its time (torchrun, N >= 2): one profiled step on rank 0 through torch.profiler:
when each NCCL kernel starts / ends relative to the compute kernels, how much of the exchange is exposed after the last
compute kernel, and how much slower the compute kernels run while NCCL kernels share the GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/ddp_timeline.py
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
    from toist_b200.models import build_model
    from toist_b200.synth import make_args, make_batch, targets_to
    from toist_b200.util.dist import DistributedDataParallel
    from toist_b200.util.misc import NestedTensor
    from toist_b200.util.optim import FusedAdamW

    torch.manual_seed(0)
    model, criterion, _, wd = build_model(make_args("resnet101", dropout=0.1))
    model.to(dev).train()
    model.enable_cuda_graphs(True)
    criterion.enable_cuda_graphs(True)
    criterion.enable_fused_loss_sum(True)
    net = DistributedDataParallel(model, device_ids=[local], find_unused_parameters=True)
    opt = FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4)
    images, mask, captions, targets, pm = make_batch(8, 640, 16, seed=1234 + rank)
    s = NestedTensor(images.to(dev), mask.to(dev))
    tg, pmd = targets_to(targets, dev), pm.to(dev)

    def step():
        mc = net(s, captions, encode_and_save=True)
        out = net(s, captions, encode_and_save=False, memory_cache=mc)
        losses = criterion(mc, out, tg, pmd, None)
        total = sum(losses[k] * wd[k] for k in losses if k in wd)
        opt.zero_grad()
        total.backward()

    for _ in range(6):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print(f"[ddp_timeline] gradients exchanged outside the stage arenas so far: {net.grad_sync.n_extra}")
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        evs.sort(key=lambda e: e.time_range.start)
        t0 = evs[0].time_range.start
        nccl = [e for e in evs if "nccl" in e.name.lower()]
        comp = [e for e in evs if "nccl" not in e.name.lower()]
        end_comp = max(e.time_range.end for e in comp) - t0
        end_all = max(e.time_range.end for e in evs) - t0
        print(f"[ddp_timeline] N={world}: step span {end_all / 1e3:.3f} ms; last compute kernel ends at {end_comp / 1e3:.3f} ms; "
              f"exposed exchange {max(end_all - end_comp, 0) / 1e3:.3f} ms; {len(nccl)} NCCL kernels, "
              f"{sum(e.time_range.end - e.time_range.start for e in nccl) / 1e3:.3f} ms of NCCL kernel time")
        for e in nccl:
            print(f"   nccl  start {(e.time_range.start - t0) / 1e3:8.3f} ms  dur {(e.time_range.end - e.time_range.start) / 1e3:7.3f} ms  {e.name[:70]}")
        # the big compute phases by first / last kernel of the known families
        def span(pred):
            xs = [e for e in comp if pred(e.name)]
            return (min(e.time_range.start for e in xs) - t0, max(e.time_range.end for e in xs) - t0) if xs else (0, 0)
        a, b = span(lambda n: "gemm_kernel" in n)
        print(f"   gemm kernels span {a / 1e3:.3f} .. {b / 1e3:.3f} ms")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
