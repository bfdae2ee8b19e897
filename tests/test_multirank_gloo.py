"""N > 1 host logic on CPU: processes over the gloo backend on 127.0.0.1 run acgpathtracing_b200.multirank.MirrorGroup, the
CPU restatement of pt_launch_multi (csrc/multigpu.cu) that takes its plan and its slices from the library's own host-only
entry points (pt_multi_plan / pt_multi_slice).  A stand-in renderer supplies each launch's mean image; every rank must end
with the image a single process folds, bit for bit.  The GPU tests (tests/test_gpu_parity.py::test_launch_multi_*) run the
real thing through the C ABI."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, SPP, DEPTH = 41, 22, 2, 6      # ragged: 41 and 22 are not multiples of the 8x4 tile


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _setup(rank, world, port):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _fake_mean(frame):
    """Stand-in for a launch's mean image: any deterministic function of (pixel, frame)."""
    rng = np.random.default_rng(1000 + frame)
    x = rng.random((H, W, 4), dtype=np.float32)
    x[..., 3] = 1.0
    return x


# ---- the plan and the slices (no processes needed) --------------------------------------------------------------------------
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8, 16])
def test_plan_covers_every_frame_and_class_once_and_is_balanced(world):
    from acgpathtracing_b200 import api
    for n in (1, 2, 3, 5, 8, 9, 16, 20, 37):
        plan = api.multi_plan(n, world)
        seen = np.zeros((n, world), np.int32)
        share = np.zeros(world, np.int64)                          # work per rank in units of 1 / world of a frame
        for frame, rank, c0, nc in plan:
            assert 0 <= frame < n and 0 <= rank < world and nc >= 1 and c0 + nc <= world
            seen[frame, c0:c0 + nc] += 1
            share[rank] += nc
        assert np.all(seen == 1)
        assert np.all(share == n)                                   # n frames * world classes / world ranks, exactly
        per_rank = [sum(1 for j in plan if j[1] == r) for r in range(world)]
        assert max(per_rank) <= n // world + 2                      # whole frames + at most two leftover jobs
        rounds = n // world
        assert [j for j in plan[:rounds * world]] == [(i * world + r, r, 0, world) for i in range(rounds) for r in range(world)]


def test_plan_examples():
    from acgpathtracing_b200 import api
    # 20 launches on 8 GPUs (the driver's scaling run): two whole frames each, then four frames in halves
    plan = api.multi_plan(20, 8)
    assert plan[16:] == [(16, 0, 0, 4), (16, 1, 4, 4), (17, 2, 0, 4), (17, 3, 4, 4), (18, 4, 0, 4), (18, 5, 4, 4), (19, 6, 0, 4), (19, 7, 4, 4)]
    # one launch on 8 GPUs (the reference's per-displayed-frame usage): the classic 1/8 tile share
    assert api.multi_plan(1, 8) == [(0, r, r, 1) for r in range(8)]


@pytest.mark.parametrize("w,h", [(41, 22), (1920, 1080), (3840, 2160), (7, 3)])
def test_slices_partition_the_image(w, h):
    from acgpathtracing_b200 import api
    for world in (1, 2, 3, 8, 16):
        s = [api.multi_slice(w, h, r, world) for r in range(world)]
        assert s[0][0] == 0 and s[-1][0] + s[-1][1] == w * h
        for a, b in zip(s, s[1:]):
            assert a[0] + a[1] == b[0]


def test_bad_arguments():
    from acgpathtracing_b200 import api
    with pytest.raises(api.PtError):
        api.multi_plan(4, 0)
    with pytest.raises(api.PtError):
        api.multi_plan(4, 17)
    with pytest.raises(api.PtError):
        api.multi_slice(16, 16, 2, 2)


# ---- MirrorGroup over gloo -----------------------------------------------------------------------------------------------------
def _mirror_worker(rank, world, port, out_dir, calls):
    _setup(rank, world, port)
    try:
        from acgpathtracing_b200 import multirank as M

        def render_raw(frame, c0, nc, n, out):
            mask = M.owned_mask(W, H, c0, n, nc)
            out[mask] = _fake_mean(frame)[mask]                    # the rest keeps MirrorGroup's NaN poison

        g = M.MirrorGroup(W, H, rank, world, render_raw)
        acc = np.zeros((H, W, 4), np.float32)
        first = 0
        for k in calls:
            mine = g.launch(acc, first, k)
            assert sum(j[3] for j in mine) == k                    # this rank's share: k / world frames
            first += k
        assert not np.isnan(acc).any()
        np.save(os.path.join(out_dir, f"acc{rank}.npy"), acc)
        # max-over-ranks reduction used by bench.py for the timed region
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == 10.0 + world - 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,calls", [(2, (4,)), (2, (5,)), (3, (2,)), (3, (7,)), (3, (1,)), (2, (3, 1, 4)), (3, (4, 5))])
def test_mirror_group_equals_sequential_fold(tmp_path, world, calls):
    """Whole frames, leftover sub-shares (including shares that span two frames) and resumed calls (first frame > 0)."""
    from acgpathtracing_b200 import multirank as M
    mp.spawn(_mirror_worker, args=(world, _free_port(), str(tmp_path), calls), nprocs=world, join=True)
    want = M.fold_numpy([_fake_mean(f) for f in range(sum(calls))])
    for r in range(world):
        got = np.load(tmp_path / f"acc{r}.npy")
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def _oracle_frame0():
    import oracle_api as O
    from acgpathtracing_b200 import scenes
    port = O.load_port()
    osc = O.OracleScene(port, scenes.cornell32())
    p = O.make_params(port, W, H, SPP, DEPTH, frame=0, nee=True, importance=True)
    return osc.render(p, math_mode=1, use_bvh=True)["accum"]


def _oracle_worker(rank, world, port, out_dir):
    _setup(rank, world, port)
    try:
        from acgpathtracing_b200 import multirank as M
        full = _oracle_frame0()                  # what the GPU kernel computes for every pixel of frame 0 ...

        def render_raw(frame, c0, nc, n, out):
            mask = M.owned_mask(W, H, c0, n, nc)
            out[mask] = full[mask]               # ... of which a rank has only its tile classes

        g = M.MirrorGroup(W, H, rank, world, render_raw)
        acc = np.zeros((H, W, 4), np.float32)
        g.launch(acc, 0, 1)                      # one launch on `world` ranks = the 1 / world tile share per rank
        np.save(os.path.join(out_dir, f"o{rank}.npy"), acc)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_single_launch_tile_shares_equal_single_rank(tmp_path, world):
    mp.spawn(_oracle_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    full = _oracle_frame0()
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"o{r}.npy").view(np.uint32), full.view(np.uint32))


def test_mirror_group_single_rank_is_the_plain_fold():
    from acgpathtracing_b200 import multirank as M
    g = M.MirrorGroup(W, H, 0, 1, lambda frame, c0, nc, n, out: out.__setitem__(Ellipsis, _fake_mean(frame)))
    acc = np.zeros((H, W, 4), np.float32)
    g.launch(acc, 0, 3)
    assert np.array_equal(acc.view(np.uint32), M.fold_numpy([_fake_mean(f) for f in range(3)]).view(np.uint32))


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
