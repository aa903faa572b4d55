#!/usr/bin/env python3
"""Aggregate host -> device feed rate of one box: every rank copies 64-frame KITTI batches (29.9 MB) to its GPU at the same time, from
(a) ordinary page-locked memory, (b) write-combined page-locked memory, (c) ordinary page-locked memory while 7.7 MB of results
travel back per batch on a second stream.  torchrun --nproc-per-node N scripts/probe/h2d_feed.py"""
import ctypes as C, glob, os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    dist.init_process_group("gloo")
torch.cuda.set_device(local)
torch.zeros(1, device="cuda")
cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + glob.glob("/usr/local/cuda/lib64/libcudart.so*")
rt = C.CDLL(sorted(cands)[0])
rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
rt.cudaStreamCreate.argtypes = [C.POINTER(C.c_void_p)]
rt.cudaStreamSynchronize.argtypes = [C.c_void_p]
N_IN, N_OUT, REP = 64 * 1241 * 376, 64 * 2000 * 60, 40
def halloc(n, flags):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), n, flags) == 0
    C.memset(p, 7, n)
    return p
d_in, d_out = C.c_void_p(), C.c_void_p()
assert rt.cudaMalloc(C.byref(d_in), N_IN) == 0 and rt.cudaMalloc(C.byref(d_out), N_OUT) == 0
s1, s2 = C.c_void_p(), C.c_void_p()
rt.cudaStreamCreate(C.byref(s1)); rt.cudaStreamCreate(C.byref(s2))
plain, wc, back = halloc(N_IN, 1), halloc(N_IN, 1 | 4), halloc(N_OUT, 1)
def run(src, with_d2h):
    if world > 1: dist.barrier()
    rt.cudaStreamSynchronize(s1); rt.cudaStreamSynchronize(s2)
    t0 = time.perf_counter()
    for _ in range(REP):
        rt.cudaMemcpyAsync(d_in, src, N_IN, 1, s1)
        if with_d2h: rt.cudaMemcpyAsync(back, d_out, N_OUT, 2, s2)
    rt.cudaStreamSynchronize(s1); rt.cudaStreamSynchronize(s2)
    dt = time.perf_counter() - t0
    t = torch.tensor([dt])
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
for name, src, d2h in (("page-locked", plain, False), ("write-combined", wc, False), ("page-locked + D2H of the results", plain, True), ("write-combined + D2H", wc, True)):
    run(src, d2h)
    dt = run(src, d2h)
    if rank == 0:
        print(f"{world} rank(s), {name:34s}: {N_IN * REP / dt / 1e9:6.1f} GB/s H2D per GPU, {world * N_IN * REP / dt / 1e9:7.1f} GB/s aggregate "
              f"-> floor {world * 64 * REP / dt / 1e3:7.1f} k frames/s", flush=True)
