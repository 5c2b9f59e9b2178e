"""CPU, world_size 2 over gloo: the ray-sharding + all-gather plumbing (panopticnerf_b200/parallel.py)
reproduces the single-process render bit for bit.  The per-rank render function here is the CPU oracle
(test infrastructure) - the product's Renderer needs a GPU; the sharding code is renderer-agnostic."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from panopticnerf_b200 import make_cfg, parallel, synthetic as S


def test_shard_range_covers_all_rays():
    for R in (0, 1, 7, 128, 529408):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(R, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == R
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            assert max(b - a for a, b in spans) == (R + world - 1) // world if R else True


def _worker(rank, world, port, R_rows, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import reference_renderer as O
    torch.set_num_threads(1)
    cfg = make_cfg("cfg1", num_classes=3, num_instances=2)
    net = S.init_network_weights(O.make_network(cfg))
    batch = S.make_batch(cfg, rows=R_rows, row0=10, num_boxes=16)
    batch["rays"] = batch["rays"][: batch["rays"].shape[0] - 3]          # not divisible by world
    ren = O.make_renderer(cfg, net)
    keys = ("rgb_map", "depth_map", "acc_map", "semantic_map")
    got = parallel.render_sharded(ren.render, batch, keys)
    if rank == 0:
        full = ren.render(batch)
        ok = all(torch.equal(got[k], full[k]) for k in keys)
        out_q.put((ok, {k: tuple(got[k].shape) for k in keys}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_render_sharded_matches_single_process(world):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 2, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, shapes = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok, "sharded render differs from the single-process render"
    assert shapes["rgb_map"] == (2 * 64 - 3, 3)
