"""One process per GPU (run under torchrun by tests/test_gpu_multi.py::test_process_per_gpu_under_torchrun):

    python -m torch.distributed.run --nproc-per-node N tests/multi_rank_check.py

Every rank builds its replica of the scene, joins a pt_group (NCCL inside libptb200.so) and calls pt_launch_multi /
pt_launch_multi_host; every rank's image must equal the sequential pt_launch loop it renders itself on its own GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from acgpathtracing_b200 import api, scenes
    from acgpathtracing_b200.multirank import rank_group

    arr = scenes.cornell32()
    w, h, spp, md = 322, 182, 3, 16
    stream = torch.cuda.current_stream().cuda_stream
    group = rank_group(rank, world)
    info = group.info()
    assert info == dict(n_ranks=world, n_local=1, first_rank=rank, transport="nccl"), info
    gs = group.create_scene(arr)
    single = api.Scene(arr)
    try:
        for first, frames in ((0, 1), (0, world + 1), (0, 2 * world + 3), (3, world)):
            want = torch.zeros((h, w, 4), dtype=torch.float32, device=dev)
            want_fb = torch.zeros((h, w, 4), dtype=torch.uint8, device=dev)
            for f in range(first + frames):
                p = api.make_params(w, h, spp, md, frame=f, nee=True, importance=True, scene=single)
                p.accumulationBuffer, p.frameBuffer = want.data_ptr(), want_fb.data_ptr()
                api.launch(p, stream)
                if f == first - 1:
                    start = want.clone()
            acc = start.clone() if first else torch.zeros_like(want)
            fb = torch.zeros_like(want_fb)
            p = api.make_params(w, h, spp, md, frame=first, nee=True, importance=True, scene=gs)
            p.accumulationBuffer, p.frameBuffer = acc.data_ptr(), fb.data_ptr()
            group.launch(p, frames, stream)
            torch.cuda.synchronize()
            assert torch.equal(acc.view(torch.int32), want.view(torch.int32)), (rank, first, frames)
            assert torch.equal(fb, want_fb), (rank, first, frames)
            # host buffers
            h_acc = start.cpu().numpy().copy() if first else np.zeros((h, w, 4), np.float32)
            h_fb = np.zeros((h, w, 4), np.uint8)
            p.accumulationBuffer, p.frameBuffer = h_acc.ctypes.data, h_fb.ctypes.data
            up, down = group.launch_host(p, frames)
            assert down == w * h * 20
            assert np.array_equal(h_acc.view(np.uint32), want.cpu().numpy().view(np.uint32)), (rank, first, frames, "host")
            assert np.array_equal(h_fb, want_fb.cpu().numpy())
        t = group.last_timing(0)
        assert t["kernels"] >= 1
        dist.barrier()
        if rank == 0:
            print(f"multi_rank_check ok: {world} ranks, timing of the last call {t}")
    finally:
        single.destroy()
        group.destroy()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
