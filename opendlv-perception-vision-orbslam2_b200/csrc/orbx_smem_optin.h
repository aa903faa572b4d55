// orbx_smem_optin.h -- opt-in to more than the default dynamic shared memory, safe across handles and host threads.
//
// cudaFuncAttributeMaxDynamicSharedMemorySize is state of the (kernel, device) pair, shared by every handle and every host
// thread of the process.  Setting it to "what this launch needs" races: another thread with a smaller geometry may lower it
// between this thread's set and its launch, which then fails with cudaErrorInvalidValue.  So the limit is only ever RAISED,
// under a lock.  The opt-in starts at 32 KB: the 48 KB default covers static + dynamic shared memory together.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <utility>

inline cudaError_t orbx_raise_dyn_smem(const void *func, size_t bytes)
{
    if (bytes <= 32 * 1024) return cudaSuccess;
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, size_t> limit;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mu);
    size_t &cur = limit[std::make_pair(func, dev)];
    if (bytes <= cur) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) cur = bytes;
    return e;
}
