"""CPU: the multi-rank host logic (sharding, index broadcast, result gather) with world_size 2 over gloo."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from effocr_b200 import dist as D
    r, w, _ = D.init_from_env("gloo")
    vec = torch.arange(12, dtype=torch.float32).reshape(4, 3) if r == 0 else None
    chars = list("abcd") if r == 0 else None
    v, c = D.broadcast_index(vec, chars)
    mine = D.shard_indices(11, r, w)
    local = {f"line{i:02d}": f"text{i}" for i in mine}
    merged = D.gather_results(local)
    q.put((r, v.tolist(), c, mine, merged))
    torch.distributed.destroy_process_group()


def test_two_rank_broadcast_shard_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs])
    for p in procs:
        p.join(60)
    (r0, v0, c0, m0, g0), (r1, v1, c1, m1, g1) = res
    assert v0 == v1 == torch.arange(12, dtype=torch.float32).reshape(4, 3).tolist()
    assert c0 == c1 == list("abcd")
    assert sorted(m0 + m1) == list(range(11)) and not set(m0) & set(m1)
    assert g1 is None and list(g0) == [f"line{i:02d}" for i in range(11)]
    assert g0["line07"] == "text7"


def test_shard_indices_weighted_is_partition_and_balanced():
    from effocr_b200 import dist as D
    w = [30, 5, 22, 41, 9, 17, 33, 28, 12, 40, 7]
    parts = [D.shard_indices(len(w), r, 4, weights=w) for r in range(4)]
    assert sorted(sum(parts, [])) == list(range(len(w)))
    loads = [sum(w[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(w)
    # results independent of world size: the union of shards is always the full set
    for world in (1, 2, 3, 8):
        assert sorted(sum((D.shard_indices(len(w), r, world) for r in range(world)), [])) == list(range(len(w)))


def test_run_effocr_sharded_single_process_uses_pipeline_contract():
    """The sharded driver only needs `infer_lines`; with world size 1 it returns every key, sorted."""
    from effocr_b200.infer import run_effocr_sharded

    class Fake:
        def infer_lines(self, imgs):
            return [{"text": f"t{int(im[0, 0, 0])}"} for im in imgs]

    import numpy as np
    imgs = [np.full((2, 2, 3), i, np.uint8) for i in range(7)]
    out = run_effocr_sharded(imgs, Fake(), keys=[f"k{i}" for i in range(7)], batch_lines=3)
    assert out == {f"k{i}": f"t{i}" for i in range(7)}
