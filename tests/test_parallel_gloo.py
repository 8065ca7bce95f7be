"""world_size-2 (and 3, ragged) CPU test of the shard + allgather host logic over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _fake_recognize(image, overlap, scene, ious):
    """Deterministic per-crop stand-in for Model.recognize (the kernels need a GPU)."""
    b = image.shape[0]
    key = (image.reshape(b, -1).sum(1) * 1000).to(torch.int64)
    ids = ((key[:, None] + torch.arange(26)[None, :] * 7) % 97).to(torch.int32)
    lengths = (key % 27).to(torch.int32)
    return {"ids": ids, "lengths": lengths, "logits": torch.zeros(b, 26, 97)}


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from multimodal_scene_text_recognition_b200 import parallel
    r, lr, w = parallel.init_from_env("gloo")
    g = torch.Generator().manual_seed(5)
    img = torch.rand(B, 1, 32, 100, generator=g)
    ov, sc, io = torch.zeros(B, 15), torch.zeros(B, 52), torch.zeros(B, 52)
    res = parallel.recognize_sharded(_fake_recognize, img, ov, sc, io, r, w)
    full = _fake_recognize(img, ov, sc, io)
    ok = torch.equal(res["ids"], full["ids"]) and torch.equal(res["lengths"], full["lengths"])
    lo, hi = res["bounds"]
    ok = ok and torch.equal(res["local_ids"], full["ids"][lo:hi])
    # the overlapped form: two exchanges in flight, waited for in order
    h1 = parallel.allgather_ids_async(res["local_ids"], res["local_lengths"], B)
    h2 = parallel.allgather_ids_async(res["local_ids"] + 1, res["local_lengths"], B)
    i1, l1 = h1.wait()
    i2, _ = h2.wait()
    ok = ok and torch.equal(i1, full["ids"]) and torch.equal(l1, full["lengths"]) and torch.equal(i2, full["ids"] + 1)
    q.put((rank, bool(ok), (lo, hi)))
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,B", [(2, 64), (3, 10), (2, 1)])
def test_sharded_recognition_gathers_global_ids(world, B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in results), results
    spans = sorted(b for _, _, b in results)
    assert spans[0][0] == 0 and spans[-1][1] == B
