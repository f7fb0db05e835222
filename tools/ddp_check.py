"""2-rank check of toist_b200.util.dist.DistributedDataParallel (run under torchrun on 2 GPUs):
gradients after a synchronised backward are bit-identical on both ranks and equal the mean of the ranks' local
gradients (taken under no_sync()), for every parameter that receives a gradient.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/ddp_check.py
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
    dist.init_process_group("nccl", device_id=dev)
    from toist_b200.models import build_model
    from toist_b200.synth import make_args, make_batch, targets_to
    from toist_b200.util.dist import DistributedDataParallel
    from toist_b200.util.misc import NestedTensor

    torch.manual_seed(rank)  # different initial weights per rank: the wrapper must broadcast rank 0's
    model, criterion, _, wd = build_model(make_args("resnet50", dropout=0.0))
    model.to(dev).eval()  # no dropout anywhere (RoBERTa keeps its own 0.1 in train mode): two passes are comparable
    graphs = os.environ.get("DDP_CHECK_GRAPHS", "1") != "0"
    if graphs:
        model.enable_cuda_graphs(True)
        criterion.enable_cuda_graphs(True)
    net = DistributedDataParallel(model, device_ids=[local], find_unused_parameters=True)
    w0 = model.class_embed.weight.detach().clone()
    ws = [torch.empty_like(w0) for _ in range(world)]
    dist.all_gather(ws, w0)
    assert all(torch.equal(ws[0], w) for w in ws), "parameters were not broadcast from rank 0"
    images, mask, captions, targets, pm = make_batch(2, 224, 12, seed=100 + rank)
    if rank % 2:  # different token ids on different ranks: the sparse word-embedding exchange has something to merge
        captions = [c.replace("sit", "dig").replace("open", "lift") for c in captions[::-1]]
    s = NestedTensor(images.to(dev), mask.to(dev))
    tg, pmd = targets_to(targets, dev), pm.to(dev)

    def step():
        model.zero_grad(set_to_none=True)
        mc = net(s, captions, encode_and_save=True)
        out = net(s, captions, encode_and_save=False, memory_cache=mc)
        losses = criterion(mc, out, tg, pmd, None)
        total = sum(losses[k] * wd[k] for k in losses if k in wd)
        total.backward()
        torch.cuda.synchronize()
        return {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}

    for it in range(2):  # second pass replays the captured graphs
        with net.no_sync():
            local_g = step()
        synced = step()
        assert local_g.keys() == synced.keys()
        worst = 0.0
        devs = []
        for n in sorted(synced):
            parts = [torch.empty_like(local_g[n]) for _ in range(world)]
            dist.all_gather(parts, local_g[n])
            mean = torch.stack(parts).mean(0)
            other = [torch.empty_like(synced[n]) for _ in range(world)]
            dist.all_gather(other, synced[n])
            assert all(torch.equal(other[0], o) for o in other), f"{n}: ranks disagree after the all-reduce"
            den = float(mean.abs().max())
            if den > 0:
                dv = float((synced[n] - mean).abs().max()) / den
                devs.append((dv, n, float((synced[n] - parts[rank]).abs().max()) / den))
                worst = max(worst, dv)
        devs.sort(reverse=True)
        if rank == 0:
            print("[ddp_check] largest deviations (vs mean, name, vs own local):", devs[:6], "median", devs[len(devs) // 2][0],
                  flush=True)
        # the matcher / atomics make two backward passes of the same batch agree to ~1e-3, not bit for bit
        assert worst < 2e-2, worst
        if rank == 0:
            print(f"[ddp_check] pass {it} graphs={graphs}: {len(synced)} gradients identical across ranks, "
                  f"max deviation from the mean of local gradients {worst:.2e}", flush=True)
    # gradient accumulation: a no_sync() backward followed by a synchronised one WITHOUT zero_grad in between must leave
    # the mean over ranks of (g_a + g_b) on every rank (what torch's wrapper exchanges after no_sync steps)
    images2, mask2, _, targets2, pm2 = make_batch(2, 224, 12, seed=300 + rank)
    s2, tg2, pmd2 = NestedTensor(images2.to(dev), mask2.to(dev)), targets_to(targets2, dev), pm2.to(dev)

    def backward(samples, tgs, pmap):
        mc = net(samples, captions, encode_and_save=True)
        out = net(samples, captions, encode_and_save=False, memory_cache=mc)
        losses = criterion(mc, out, tgs, pmap, None)
        sum(losses[k] * wd[k] for k in losses if k in wd).backward()
        torch.cuda.synchronize()

    def grads():
        return {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}

    with net.no_sync():
        model.zero_grad(set_to_none=True)
        backward(s, tg, pmd)
        ga = grads()
        model.zero_grad(set_to_none=True)
        backward(s2, tg2, pmd2)
        gb = grads()
        model.zero_grad(set_to_none=True)
        backward(s, tg, pmd)  # accumulated locally ...
    backward(s2, tg2, pmd2)  # ... and exchanged together with this pass
    acc = grads()
    worst = 0.0
    for n in sorted(acc):
        want = ga[n] + gb[n]
        dist.all_reduce(want)
        want /= world
        other = [torch.empty_like(acc[n]) for _ in range(world)]
        dist.all_gather(other, acc[n])
        assert all(torch.equal(other[0], o) for o in other), f"{n}: ranks disagree after the accumulated exchange"
        den = float(want.abs().max())
        if den > 0:
            worst = max(worst, float((acc[n] - want).abs().max()) / den)
    assert worst < 2e-2, worst
    if rank == 0:
        print(f"[ddp_check] no_sync + sync accumulation: {len(acc)} gradients identical across ranks, max deviation from "
              f"the mean of the summed local gradients {worst:.2e}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
