"""N > 1 host path on CPU: two processes, gloo backend, 127.0.0.1.  Each rank "renders" only the pixels it owns (the
CPU oracle stands in for the GPU kernel, launch by launch, keeping its own running average), then the ranks run the
product's exchange (acgpathtracing_b200.multirank.ImageExchange: pack -> all_gather -> unpack) ONCE after the last
launch.  Both ranks must end with the image a single process renders, bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, SPP, DEPTH, FRAMES = 41, 22, 2, 6, 3      # ragged: 41 and 22 are not multiples of the 8x4 tile


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _render_single():
    import oracle_api as O
    from acgpathtracing_b200 import scenes
    port = O.load_port()
    osc = O.OracleScene(port, scenes.cornell32())
    acc = None
    for f in range(FRAMES):
        p = O.make_params(port, W, H, SPP, DEPTH, frame=f, nee=True, importance=True)
        o = osc.render(p, accum=acc, math_mode=1, use_bvh=True)
        acc = o["accum"]
    return acc, o["fb"]


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from acgpathtracing_b200 import multirank as M
        full_acc, full_fb = _render_single()          # what the GPU kernel would compute for every pixel ...
        mine = M.owned_mask(W, H, rank, world)
        acc = torch.zeros((H, W, 4), dtype=torch.float32)
        fb = torch.zeros((H, W, 4), dtype=torch.uint8)
        acc.numpy()[mine] = full_acc[mine]            # ... of which this rank has only its own tiles
        fb.numpy()[mine] = full_fb[mine]
        assert not np.array_equal(acc.numpy(), full_acc)
        ex = M.ImageExchange(W, H, rank, world, "cpu")
        ex.gather(acc, fb)
        np.save(os.path.join(out_dir, f"acc{rank}.npy"), acc.numpy())
        np.save(os.path.join(out_dir, f"fb{rank}.npy"), fb.numpy())
        # max-over-ranks reduction used by bench.py for the timed region
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == 10.0 + world - 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_rank_exchange_equals_single_rank(tmp_path, world):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    full_acc, full_fb = _render_single()
    for r in range(world):
        a = np.load(tmp_path / f"acc{r}.npy")
        f = np.load(tmp_path / f"fb{r}.npy")
        assert np.array_equal(a.view(np.uint32), full_acc.view(np.uint32))
        assert np.array_equal(f, full_fb)


def test_exchange_is_identity_for_one_rank():
    from acgpathtracing_b200 import multirank as M
    ex = M.ImageExchange(16, 8, 0, 1, "cpu")
    a = torch.ones((8, 16, 4))
    assert ex.gather(a, torch.zeros((8, 16, 4), dtype=torch.uint8)) == 0 and ex.bytes_per_rank() == 0
