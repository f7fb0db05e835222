"""world_size-2 gloo tests (CPU) of the host-side multi-rank logic: the global num_boxes normalisation of the
criterion (reference models/mdetr.py:997-1001), per-rank synthetic shards, and the bench reference arm on rank != 0."""
from __future__ import annotations

import os
import socket
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank: int, world: int, port: int, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from toist_b200.models.matcher import HungarianMatcher
    from toist_b200.models.mdetr import SetCriterion
    from toist_b200.synth import make_args, make_batch

    crit = SetCriterion(make_args(), 255, HungarianMatcher(1, 5, 2), 0.1, ["labels", "boxes", "cardinality"], 0.07, 64)
    # rank 0 has 3 images (1 + 2 + 3 boxes), rank 1 has 1 image (1 box): global mean (6 + 1) / 2 = 3.5
    _, _, _, targets, _ = make_batch(3 if rank == 0 else 1, 32, 8, seed=1234 + rank)
    nb = crit.num_boxes_tensor(targets, "cpu")
    # an all-empty step clamps to 1 (models/mdetr.py:1001)
    nb0 = crit.num_boxes_tensor([{"labels": torch.zeros(0)}], "cpu")
    # different ranks draw different synthetic shards
    img = make_batch(1, 8, 8, seed=1234 + rank)[0]
    gathered = [torch.zeros_like(img) for _ in range(world)]
    dist.all_gather(gathered, img)
    # flat gradient exchange (util/dist.py): a stage's arena and one gradient living outside of it
    from toist_b200.util.dist import FlatGradSync

    arena = torch.arange(8, dtype=torch.float32) * (rank + 1)
    views = (arena[:4].view(2, 2), None, arena[4:8])
    stray = torch.full((3,), float(rank + 1))
    sync = FlatGradSync()
    sync.reduce("stage", arena, views + (stray,))
    sync.enabled = False
    untouched = torch.ones(2) * (rank + 1)
    sync.reduce("stage", untouched, (untouched,))
    flat_ok = bool(torch.equal(arena, torch.arange(8, dtype=torch.float32) * 1.5) and torch.equal(stray, torch.full((3,), 1.5))
                   and torch.equal(views[0], arena[:4].view(2, 2)) and torch.equal(untouched, torch.ones(2) * (rank + 1)))
    import argparse

    import bench

    out = bench.run_reference(argparse.Namespace(gpus=2, steps=1, warmup=1)) if rank != 0 else "skipped-on-rank0"
    q.put((rank, float(nb), float(nb0), bool(torch.equal(gathered[0], gathered[1])), out, flat_ok))
    dist.destroy_process_group()


def test_two_rank_gloo_host_logic():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, nb, nb0, same_shard, ref, flat_ok in res:
        assert flat_ok  # averaged in place, views intact, no_sync() leaves gradients alone
        assert nb == 3.5
        assert nb0 == 1.0
        assert not same_shard
    assert res[1][4] is None  # the reference arm does no work on rank != 0


def test_memory_cache_survives_ddp_argument_mapping():
    """DistributedDataParallel maps every (nested) dict among the forward arguments through `type(obj)(items)`
    (torch/distributed/utils.py:_recursive_to).  `memory_cache` is such an argument in phase B (engine.py:66) and
    carries the tokenizer output: it must come out with its caption lengths intact (char_to_token feeds the
    contrastive-alignment loss, models/mdetr.py:614-645)."""
    import torch
    from torch.distributed.utils import _to_kwargs

    from toist_b200.tokenizer import CharTokenizer

    tok = CharTokenizer()(["open something", "cut"])
    mc = {"tokenized": tok, "mask": torch.zeros(2, 4, dtype=torch.bool), "nested": {"t": tok}}
    _, kw = _to_kwargs((), {"memory_cache": mc, "encode_and_save": False}, torch.device("cpu"), False)
    out = kw[0]["memory_cache"]
    for t in (out["tokenized"], out["nested"]["t"]):
        assert t.char_to_token(0, 3) == 4 and t.char_to_token(1, 2) == 3 and t.char_to_token(1, 3) is None
        assert torch.equal(t["input_ids"], tok["input_ids"]) and torch.equal(t.attention_mask, tok.attention_mask)
