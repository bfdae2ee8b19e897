"""Owned-pixel exchange between ranks (one process per GPU): the one collective of the multi-GPU path.

Each rank renders the 8x4-pixel tiles it owns (pt_launch_partition) into full-size buffers for every launch of an
image, keeps their running average on its own device, and after the LAST launch calls ``ImageExchange.gather`` once:
pack owned pixels -> all-gather (NCCL on GPUs, gloo in the CPU tests) -> unpack, after which every rank holds the
complete accumulation and frame buffers, bit-identical to a single-GPU run.

The pack/unpack index math is acgpathtracing_b200/partition.py; on CUDA tensors it runs as the library's
partition_pack/unpack kernels (include/ptb200.h), on CPU tensors as the numpy mirror -- same layout, same result.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import partition as P


class ImageExchange:
    def __init__(self, width, height, rank, world, device):
        self.w, self.h, self.rank, self.world = width, height, rank, world
        self.device = torch.device(device)
        self.cap = P.capacity(width, height, world)
        self.cuda = self.device.type == "cuda"
        if world > 1:
            self.my_acc = torch.zeros((self.cap, 4), dtype=torch.float32, device=self.device)
            self.my_fb = torch.zeros((self.cap, 4), dtype=torch.uint8, device=self.device)
            # rank-major concatenation (the layout pt_partition_unpack reads)
            self.all_acc = torch.zeros((world * self.cap, 4), dtype=torch.float32, device=self.device)
            self.all_fb = torch.zeros((world * self.cap, 4), dtype=torch.uint8, device=self.device)

    def bytes_per_rank(self):
        """Bytes this rank receives in one gather (accumulation + frame buffer of all ranks)."""
        return 0 if self.world == 1 else self.world * self.cap * 20

    def gather(self, acc, fb, stream=None):
        """acc: (h, w, 4) float32, fb: (h, w, 4) uint8 or None -- full-size, valid at this rank's pixels; completed in
        place.  Returns the number of library kernels launched."""
        if self.world == 1:
            return 0
        if self.cuda:
            from . import api
            api.partition_pack(acc.data_ptr(), fb.data_ptr() if fb is not None else 0, self.my_acc.data_ptr(),
                               self.my_fb.data_ptr() if fb is not None else 0, self.w, self.h, self.rank, self.world, stream)
        else:
            self.my_acc.copy_(torch.from_numpy(P.pack_numpy(acc.numpy(), self.w, self.h, self.rank, self.world)))
            if fb is not None:
                self.my_fb.copy_(torch.from_numpy(P.pack_numpy(fb.numpy(), self.w, self.h, self.rank, self.world)))
        dist.all_gather_into_tensor(self.all_acc, self.my_acc)
        if fb is not None:
            dist.all_gather_into_tensor(self.all_fb, self.my_fb)
        if self.cuda:
            from . import api
            api.partition_unpack(self.all_acc.data_ptr(), self.all_fb.data_ptr() if fb is not None else 0, acc.data_ptr(),
                                 fb.data_ptr() if fb is not None else 0, self.w, self.h, self.world, stream)
        else:
            acc.copy_(torch.from_numpy(P.unpack_numpy(self.all_acc.numpy().reshape(self.world, self.cap, 4), self.w, self.h, self.world)))
            if fb is not None:
                fb.copy_(torch.from_numpy(P.unpack_numpy(self.all_fb.numpy().reshape(self.world, self.cap, 4), self.w, self.h, self.world)))
        return 2


def owned_mask(width, height, rank, world):
    """bool[h, w]: pixels rendered by `rank`."""
    return P.owner_map(width, height, world) == rank


# ---- frame-parallel rendering: several launches ("frames") per image, more than one rank ---------------------------------
def frame_plan(n_frames, world):
    """Which rank renders what.  The mean image of frame k depends only on (pixel, k) (pathTracerPrograms.cu:721), so
    whole frames can go to different ranks: frames 0 .. rounds*world-1 are dealt round-robin (frame k -> rank k % world,
    full image, no tail of a small per-rank pixel share); the n_frames % world leftover frames are rendered by all ranks
    together with the tile partition.  Returns (rounds, leftover_frames)."""
    rounds = n_frames // world
    return rounds, list(range(rounds * world, n_frames))


def fold_numpy(means, first_frame=0, acc=None):
    """CPU mirror of pt_fold_frames: the reference's running average (pathTracerPrograms.cu:803-811) over per-launch mean
    images (h, w, 4) float32, in frame order.  Returns (h, w, 4) with alpha 1."""
    out = None if acc is None else acc[..., :3].astype(np.float32).copy()
    for k, x in enumerate(means):
        f = first_frame + k
        x3 = x[..., :3].astype(np.float32)
        if f == 0:
            out = x3.copy()
        else:
            a = np.float32(1.0) / np.float32(f + 1)
            out = out + a * (x3 - out)
    res = np.ones(out.shape[:2] + (4,), np.float32)
    res[..., :3] = out
    return res


class FrameParallel:
    """Renders `n_frames` launches of one image on `world` ranks and leaves the complete running average (and frame
    buffer) on every rank, bit-identical to the same launches issued one after the other on one GPU.

    render_raw(frame, rank, n_ranks, out) must write the launch's own mean for the tiles of (rank, n_ranks) into the
    (h, w, 4) float32 tensor `out` (pt_launch_raw on GPUs; the CPU tests pass a stand-in)."""

    def __init__(self, width, height, rank, world, device, n_frames, render_raw, fold=None, make_color=None):
        self.w, self.h, self.rank, self.world, self.n = width, height, rank, world, n_frames
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.render_raw, self.fold_fn, self.make_color = render_raw, fold, make_color
        self.rounds, self.leftover = frame_plan(n_frames, world)
        f32 = dict(dtype=torch.float32, device=self.device)
        self.mine = torch.zeros((max(self.rounds, 1), height, width, 4), **f32)
        self.all_whole = torch.zeros((world * max(self.rounds, 1), height, width, 4), **f32) if world > 1 else self.mine
        self.shared = torch.zeros((max(len(self.leftover), 1), height, width, 4), **f32)
        self.exchange = ImageExchange(width, height, rank, world, device)

    def my_frames(self):
        """(frame index, kind) in the order this rank renders them."""
        return [(i * self.world + self.rank, "whole") for i in range(self.rounds)] + [(f, "tiles") for f in self.leftover]

    def render(self, frame, kind):
        if kind == "whole":
            self.render_raw(frame, 0, 1, self.mine[frame // self.world])
        else:
            self.render_raw(frame, self.rank, self.world, self.shared[frame - self.rounds * self.world])
        return 1

    def bytes_received(self):
        if self.world == 1:
            return 0
        return self.world * self.rounds * self.w * self.h * 16 + len(self.leftover) * self.world * self.exchange.cap * 16

    def finish(self, acc, fb, stream=None):
        """Exchange the means and fold them in frame order into acc (h, w, 4) f32 and fb (h, w, 4) u8 (or None)."""
        launches = 0
        if self.world > 1:
            if self.rounds:
                dist.all_gather_into_tensor(self.all_whole.view(self.world * self.rounds, -1), self.mine.view(self.rounds, -1))
            for i in range(len(self.leftover)):
                launches += self.exchange.gather(self.shared[i], None, stream)
        frames = []
        for k in range(self.n):
            if k < self.rounds * self.world:
                r, i = k % self.world, k // self.world
                frames.append(self.all_whole[r * self.rounds + i] if self.world > 1 else self.mine[i])
            else:
                frames.append(self.shared[k - self.rounds * self.world])
        if self.cuda:
            from . import api
            for c in range(0, self.n, 64):
                api.fold_frames(acc.data_ptr(), fb.data_ptr() if fb is not None else 0, [t.data_ptr() for t in frames[c:c + 64]],
                                c, self.w, self.h, stream)
                launches += 1
        else:
            res = fold_numpy([t.numpy() for t in frames])
            acc.copy_(torch.from_numpy(res))
            if fb is not None and self.make_color is not None:
                fb.copy_(torch.from_numpy(self.make_color(res)))
        return launches
