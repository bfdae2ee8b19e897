"""Multi-GPU rendering with one process per GPU (torchrun): the host side.

The product path is ``rank_group`` -> ``api.Group.launch``: plan, slice exchange (NCCL), ordered fold and gather all happen
inside ``libptb200.so`` (csrc/multigpu.cu, include/ptb200.h ``pt_launch_multi``); Python only ships the 128-byte NCCL id
from rank 0 to the other ranks.

``MirrorGroup`` restates the same algorithm on CPU tensors -- the plan and the slices come from the library's own host-only
entry points (``pt_multi_plan``, ``pt_multi_slice``), the exchange runs over ``torch.distributed`` (gloo in the CPU tests),
the fold is numpy -- so that the N > 1 logic (who renders what, which piece of a frame a pixel is read from, the frame
order of the running average, pathTracerPrograms.cu:803-811) is tested with world_size 2 and 3 on machines without a GPU
(tests/test_multirank_gloo.py).  It is test infrastructure: nothing on the GPU path calls it.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import partition as P


def rank_group(rank, world, device_index=None):
    """One process per GPU: every rank calls this once (collective).  Returns an api.Group on the current CUDA device."""
    from . import api
    if device_index is not None:
        torch.cuda.set_device(device_index)
    box = [api.group_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    return api.Group.rank(box[0], rank, world)


def owned_mask(width, height, rank, world, n_classes=1):
    """bool[h, w]: pixels of the tiles whose class (tile index % world) is in [rank, rank + n_classes)."""
    cls = P.owner_map(width, height, world)
    return (cls >= rank) & (cls < rank + n_classes)


def fold_numpy(means, first_frame=0, acc=None):
    """CPU mirror of the fold kernels: the reference's running average (pathTracerPrograms.cu:803-811) over per-launch mean
    images (..., 4) float32, in frame order.  Returns an array of the same shape with alpha 1."""
    out = None if acc is None else acc[..., :3].astype(np.float32).copy()
    for k, x in enumerate(means):
        f = first_frame + k
        x3 = x[..., :3].astype(np.float32)
        if f == 0:
            out = x3.copy()
        else:
            a = np.float32(1.0) / np.float32(f + 1)
            out = out + a * (x3 - out)
    res = np.ones(out.shape[:-1] + (4,), np.float32)
    res[..., :3] = out
    return res


class MirrorGroup:
    """CPU restatement of pt_launch_multi for one rank of `world` (see the module docstring).

    render_raw(frame, first_class, n_classes, n_ranks, out): writes the launch's own mean of `frame` for the tiles of
    classes [first_class, first_class + n_classes) of n_ranks into the (h, w, 4) float32 array `out` (whole frame:
    first_class 0, n_classes == n_ranks).  Pixels outside the job may hold anything -- the fold never reads them."""

    def __init__(self, width, height, rank, world, render_raw):
        from . import api
        self.api = api
        self.w, self.h, self.rank, self.world = width, height, rank, world
        self.render_raw = render_raw
        self.slices = [api.multi_slice(width, height, r, world) for r in range(world)]
        self.cls = P.owner_map(width, height, world).reshape(-1)

    def launch(self, acc, first_frame, n_launches):
        """acc: (h, w, 4) float32 numpy, complete image of the frames before first_frame (ignored when first_frame == 0);
        folded in place with frames first_frame .. first_frame + n_launches - 1.  Returns the jobs this rank rendered."""
        w, h, me, world = self.w, self.h, self.rank, self.world
        if world == 1:
            means = []
            for k in range(n_launches):
                out = np.zeros((h, w, 4), np.float32)
                self.render_raw(first_frame + k, 0, 1, 1, out)
                means.append(out)
            acc[...] = fold_numpy(means, first_frame, acc)
            return [(k, 0, 0, 1) for k in range(n_launches)]
        plan = self.api.multi_plan(n_launches, world)
        rendered = {}
        for q, (frame, r, c0, nc) in enumerate(plan):
            if r != me:
                continue
            out = np.full((h, w, 4), np.nan, np.float32)          # poison: a read outside the job's tiles shows up
            self.render_raw(first_frame + frame, c0, nc, world, out)
            rendered[q] = out.reshape(-1, 4)
        # exchange: slice d of every job goes to rank d (same order on both sides of every pair)
        b, n = self.slices[me]
        pieces = {}
        ops, keep = [], []
        for q, (frame, r, c0, nc) in enumerate(plan):
            if r == me:
                pieces[q] = rendered[q][b:b + n]
                for d in range(world):
                    if d != me and self.slices[d][1]:
                        db, dn = self.slices[d]
                        t = torch.from_numpy(np.ascontiguousarray(rendered[q][db:db + dn]))
                        keep.append(t)
                        ops.append(dist.P2POp(dist.isend, t, d))
            elif n:
                t = torch.empty((n, 4), dtype=torch.float32)
                pieces[q] = t.numpy()
                ops.append(dist.P2POp(dist.irecv, t, r))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        # fold my slice in frame order; a frame rendered in shares is read from the piece whose classes own the pixel
        flat = acc.reshape(-1, 4)
        cls = self.cls[b:b + n]
        means = []
        for k in range(n_launches):
            m = np.zeros((n, 4), np.float32)
            for q, (frame, r, c0, nc) in enumerate(plan):
                if frame == k:
                    sel = (cls >= c0) & (cls < c0 + nc)
                    m[sel] = pieces[q][sel]
            means.append(m)
        if n:
            flat[b:b + n] = fold_numpy(means, first_frame, flat[b:b + n])
        # gather the folded slices on every rank
        cap = max(s[1] for s in self.slices)
        mine = torch.zeros((cap, 4), dtype=torch.float32)
        mine[:n] = torch.from_numpy(flat[b:b + n])
        everyone = [torch.zeros((cap, 4), dtype=torch.float32) for _ in range(world)]
        dist.all_gather(everyone, mine)
        for r, (rb, rn) in enumerate(self.slices):
            flat[rb:rb + rn] = everyone[r][:rn].numpy()
        return [j for j in plan if j[1] == me]
