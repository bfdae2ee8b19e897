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


# ---- frame-parallel plan: whole frames dealt round-robin + tile-partitioned leftovers, ordered fold --------------------
def _fake_mean(frame):
    """Stand-in for a launch's mean image: any deterministic function of (pixel, frame)."""
    rng = np.random.default_rng(1000 + frame)
    x = rng.random((H, W, 4), dtype=np.float32)
    x[..., 3] = 1.0
    return x


def _fp_worker(rank, world, port, out_dir, n_frames):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from acgpathtracing_b200 import multirank as M

        def render_raw(frame, r, n, out):
            mask = M.owned_mask(W, H, r, n)
            out.zero_()
            out.numpy()[mask] = _fake_mean(frame)[mask]

        fp = M.FrameParallel(W, H, rank, world, "cpu", n_frames, render_raw)
        mine = fp.my_frames()
        assert len(mine) == n_frames // world + n_frames % world
        for f, kind in mine:
            fp.render(f, kind)
        acc = torch.zeros((H, W, 4), dtype=torch.float32)
        fp.finish(acc, None)
        np.save(os.path.join(out_dir, f"fp{rank}.npy"), acc.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames", [(2, 4), (2, 5), (3, 2), (3, 7)])
def test_frame_parallel_equals_sequential_fold(tmp_path, world, n_frames):
    from acgpathtracing_b200 import multirank as M
    mp.spawn(_fp_worker, args=(world, _free_port(), str(tmp_path), n_frames), nprocs=world, join=True)
    want = M.fold_numpy([_fake_mean(f) for f in range(n_frames)])
    for r in range(world):
        got = np.load(tmp_path / f"fp{r}.npy")
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_frame_plan_covers_every_frame_once():
    from acgpathtracing_b200 import multirank as M
    for world in (1, 2, 3, 4, 8):
        for n in (1, 2, 5, 8, 9, 16):
            rounds, left = M.frame_plan(n, world)
            whole = sorted(i * world + r for r in range(world) for i in range(rounds))
            assert whole + left == list(range(n))


def test_fold_numpy_first_frame_and_weights(port, cornell):
    """Frame 0 overwrites (pathTracerPrograms.cu:803 reads the buffer only for frames > 0) and frame k enters with weight
    1/(k+1) in float32: the fold of [x, x, x] is x again wherever (x - x) is exact, and a constant offset converges as
    the reference's lerp chain does.  (The GPU tests compare pt_fold_frames with real launches bit for bit.)"""
    import oracle_api as O
    from acgpathtracing_b200 import multirank as M
    osc = O.OracleScene(port, cornell)
    w, h = 24, 16
    first = osc.render(O.make_params(port, w, h, 2, 4, frame=0, nee=True, importance=True), math_mode=1, use_bvh=True)["accum"]
    assert np.array_equal(M.fold_numpy([first]).view(np.uint32), first.view(np.uint32))
    assert np.array_equal(M.fold_numpy([first, first, first]).view(np.uint32), first.view(np.uint32))
    a = np.zeros((h, w, 4), np.float32)
    b = np.full((h, w, 4), 3.0, np.float32)
    got = M.fold_numpy([a, b, b])                    # 0 -> 0 + (3-0)/2 = 1.5 -> 1.5 + (3-1.5)/3 = 2
    assert np.all(got[..., :3] == np.float32(1.5) + np.float32(1.0) / np.float32(3) * np.float32(1.5))
    resumed = M.fold_numpy([b], first_frame=2, acc=M.fold_numpy([a, b]))
    assert np.array_equal(resumed.view(np.uint32), got.view(np.uint32))
