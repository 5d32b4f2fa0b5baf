"""Host->device copy bandwidth with ordinary pinned memory vs write-combined pinned memory (cudaHostAllocWriteCombined), all ranks at once.
usage: python -m torch.distributed.run --nproc-per-node N tools/wc_probe.py"""
import ctypes as C
import os
import time

import numpy as np
import torch
import torch.distributed as dist

rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1)); local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
rt = C.CDLL('libcudart.so.12')
NB = 320 * 720 * 540


def host_alloc(nbytes, flags):
    p = C.c_void_p()
    rc = rt.cudaHostAlloc(C.byref(p), C.c_size_t(nbytes), C.c_uint(flags))
    assert rc == 0, rc
    return p.value


def bw(ptr, steps=30):
    d = torch.empty(NB, dtype=torch.uint8, device='cuda')
    s = torch.cuda.Stream()
    rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    def run(k):
        for _ in range(k):
            rt.cudaMemcpyAsync(d.data_ptr(), ptr, NB, 1, s.cuda_stream)
        s.synchronize()
    run(3)
    best = 1e9
    for _ in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter(); run(steps); dt = time.perf_counter() - t0
        t = torch.tensor([dt], device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t))
    return NB * steps / best / 1e9


for name, flags in (('pinned', 0), ('pinned+portable', 1), ('write-combined', 4), ('write-combined+portable', 5)):
    ptr = host_alloc(NB, flags)
    a = np.ctypeslib.as_array((C.c_uint8 * NB).from_address(ptr))
    a[:] = (np.arange(NB) % 251).astype(np.uint8)           # first touch on this rank's node
    g = bw(ptr)
    if rank == 0:
        print('%-26s %6.1f GB/s per GPU, %7.1f GB/s aggregate (%d ranks)' % (name, g, g * world, world), flush=True)
    rt.cudaFreeHost(C.c_void_p(ptr))
if world > 1:
    dist.destroy_process_group()
