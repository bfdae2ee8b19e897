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
        """acc: (h, w, 4) float32, fb: (h, w, 4) uint8 -- full-size, valid at this rank's pixels; completed in place.
        Returns the number of library kernels launched."""
        if self.world == 1:
            return 0
        if self.cuda:
            from . import api
            api.partition_pack(acc.data_ptr(), fb.data_ptr(), self.my_acc.data_ptr(), self.my_fb.data_ptr(), self.w, self.h,
                               self.rank, self.world, stream)
        else:
            self.my_acc.copy_(torch.from_numpy(P.pack_numpy(acc.numpy(), self.w, self.h, self.rank, self.world)))
            self.my_fb.copy_(torch.from_numpy(P.pack_numpy(fb.numpy(), self.w, self.h, self.rank, self.world)))
        dist.all_gather_into_tensor(self.all_acc, self.my_acc)
        dist.all_gather_into_tensor(self.all_fb, self.my_fb)
        if self.cuda:
            from . import api
            api.partition_unpack(self.all_acc.data_ptr(), self.all_fb.data_ptr(), acc.data_ptr(), fb.data_ptr(), self.w, self.h,
                                 self.world, stream)
        else:
            acc.copy_(torch.from_numpy(P.unpack_numpy(self.all_acc.numpy().reshape(self.world, self.cap, 4), self.w, self.h, self.world)))
            fb.copy_(torch.from_numpy(P.unpack_numpy(self.all_fb.numpy().reshape(self.world, self.cap, 4), self.w, self.h, self.world)))
        return 2


def owned_mask(width, height, rank, world):
    """bool[h, w]: pixels rendered by `rank`."""
    return P.owner_map(width, height, world) == rank
