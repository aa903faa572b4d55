// orbm_match.cu -- Hamming kNN-2 matcher (include/orbx.h, orbm_* entry points).
//
// Reproduces ORBmatcher::DescriptorDistance (orbmatcher.cpp:1662-1677: 8 x 32-bit XOR + popcount)
// evaluated over every query x train pair, with the reference's best / second-best bookkeeping
// (orbmatcher.cpp:208-232: strict '<', start values 256 / -1; lowest index wins ties and the
// second best counts duplicates of the best).
//
// Kernel shape: INT pipe only (LOP3 + POPC + IADD3 + IMNMX) -- this is not a dense float
// contraction, tensor cores do not apply.  One query per thread held in 8 registers; the train
// set is streamed through shared memory in tiles that every thread of the CTA reads as 128-bit
// broadcasts; the grid is (query blocks) x (train chunks) so that 2000 queries still fill 148 SMs.
// The two smallest (distance << 22 | index) keys per query are kept -- keys are unique, so
// min-of-keys is "lowest index attains the minimum" and the second smallest key carries the
// second-best distance with multiplicity.  A merge kernel folds the per-chunk partials.
#include "../../include/orbx.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#define KNN_QB 128          // queries per CTA (one per thread)
#define KNN_TILE 256        // train descriptors per shared-memory tile (8 KB)
#define KNN_IDX_BITS 22
#define KNN_IDX_MASK 0x3fffffu
#define KNN_INIT_KEY ((256u << KNN_IDX_BITS) | KNN_IDX_MASK)

__global__ void __launch_bounds__(KNN_QB)
k_knn2_partial(const uint4 *__restrict__ q, int nq, const uint4 *__restrict__ t, int nt, int chunk,
               uint2 *__restrict__ part)
{
    __shared__ uint4 tile[KNN_TILE * 2];
    const int tid = threadIdx.x;
    const int qi = blockIdx.x * KNN_QB + tid;
    const int t0 = blockIdx.y * chunk;
    const int t1 = min(t0 + chunk, nt);
    uint4 qa = make_uint4(0, 0, 0, 0), qb = qa;
    if (qi < nq) { qa = __ldg(&q[2 * qi]); qb = __ldg(&q[2 * qi + 1]); }
    unsigned k1 = KNN_INIT_KEY, k2 = KNN_INIT_KEY;
    for (int base = t0; base < t1; base += KNN_TILE) {
        const int cnt = min(KNN_TILE, t1 - base);
        __syncthreads();
        for (int i = tid; i < cnt * 2; i += KNN_QB) tile[i] = __ldg(&t[2 * (size_t)base + i]);
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < cnt; j++) {
            const uint4 a = tile[2 * j], b = tile[2 * j + 1];
            const int d = __popc(qa.x ^ a.x) + __popc(qa.y ^ a.y) + __popc(qa.z ^ a.z) + __popc(qa.w ^ a.w) +
                          __popc(qb.x ^ b.x) + __popc(qb.y ^ b.y) + __popc(qb.z ^ b.z) + __popc(qb.w ^ b.w);
            const unsigned key = ((unsigned)d << KNN_IDX_BITS) | (unsigned)(base + j);
            k2 = min(k2, max(k1, key));
            k1 = min(k1, key);
        }
    }
    if (qi < nq) part[(size_t)blockIdx.y * nq + qi] = make_uint2(k1, k2);
}

__global__ void k_knn2_merge(const uint2 *__restrict__ part, int nq, int nchunks, int4 *__restrict__ out)
{
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    unsigned k1 = KNN_INIT_KEY, k2 = KNN_INIT_KEY;
    for (int c = 0; c < nchunks; c++) {
        const uint2 p = part[(size_t)c * nq + qi];
        // fold two sorted pairs: the new pair is the two smallest of {k1,k2,p.x,p.y}
        k2 = min(k2, max(k1, p.x));
        k1 = min(k1, p.x);
        k2 = min(k2, max(k1, p.y));
        k1 = min(k1, p.y);
    }
    int d1 = (int)(k1 >> KNN_IDX_BITS), d2 = (int)(k2 >> KNN_IDX_BITS);
    int idx = (int)(k1 & KNN_IDX_MASK);
    if (d1 >= 256) { d1 = 256; idx = -1; }  // 'dist < bestDist1' with bestDist1 = 256 never fires
    if (d2 >= 256) d2 = 256;
    out[qi] = make_int4(idx, d1, d2, 0);
}

__global__ void k_distance_pairs(const uint4 *__restrict__ a, const uint4 *__restrict__ b, int n, int *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 a0 = a[2 * i], a1 = a[2 * i + 1], b0 = b[2 * i], b1 = b[2 * i + 1];
    out[i] = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
             __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

struct orbm_matcher {
    int device = 0, maxQ = 0, maxT = 0, smCount = 148;
    cudaStream_t stream = nullptr;
    uint8_t *dQ = nullptr, *dT = nullptr;
    uint2 *dPart = nullptr; size_t partCap = 0;
    int4 *dOut = nullptr;
    int4 *hOut = nullptr;
    int residentNt = -1;
    std::string err;
};

namespace {
int mfail(orbm_matcher *m, int code, const std::string &msg) { if (m) m->err = msg; return code; }
#define MCK(call)                                                                                     \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) return mfail(m, ORBX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

// train chunks so that (query blocks x chunks) is a whole number of waves of the SM count
void knnGrid(const orbm_matcher *m, int nq, int nt, int *qBlocks, int *chunks, int *chunk)
{
    const int qb = (nq + KNN_QB - 1) / KNN_QB;
    int want = std::max(1, (m->smCount * 4 + qb - 1) / qb);
    int maxChunks = std::max(1, (nt + KNN_TILE - 1) / KNN_TILE);
    int c = std::min(want, maxChunks);
    int per = ((nt + c - 1) / c + KNN_TILE - 1) / KNN_TILE * KNN_TILE;
    if (per < KNN_TILE) per = KNN_TILE;
    c = std::max(1, (nt + per - 1) / per);
    *qBlocks = qb; *chunks = c; *chunk = per;
}

int enqueueKnn(orbm_matcher *m, const uint8_t *dq, int nq, const uint8_t *dt, int nt, int4 *dout, cudaStream_t st)
{
    if (((uintptr_t)dq | (uintptr_t)dt | (uintptr_t)dout) & 15) return mfail(m, ORBX_ERR_ARG, "device buffers must be 16-byte aligned");
    int qb, chunks, chunk;
    knnGrid(m, nq, std::max(nt, 1), &qb, &chunks, &chunk);
    if (nt == 0) chunks = 0;
    const size_t need = (size_t)std::max(chunks, 1) * nq;
    if (need > m->partCap) {
        if (m->dPart) cudaFree(m->dPart);
        m->dPart = nullptr; m->partCap = 0;
        MCK(cudaMalloc((void **)&m->dPart, need * sizeof(uint2)));
        m->partCap = need;
    }
    if (chunks > 0) {
        dim3 grid(qb, chunks);
        k_knn2_partial<<<grid, KNN_QB, 0, st>>>((const uint4 *)dq, nq, (const uint4 *)dt, nt, chunk, m->dPart);
    }
    k_knn2_merge<<<(nq + 127) / 128, 128, 0, st>>>(m->dPart, nq, chunks, dout);
    MCK(cudaGetLastError());
    return ORBX_OK;
}
} // namespace

extern "C" {

int orbm_create(int device, int max_queries, int max_train, orbm_matcher **out)
{
    if (!out || max_queries < 1 || max_train < 1 || max_train > (int)KNN_IDX_MASK - 1) return ORBX_ERR_ARG;
    orbm_matcher *m = new (std::nothrow) orbm_matcher();
    if (!m) return ORBX_ERR_NOMEM;
    *out = m;
    m->device = device; m->maxQ = max_queries; m->maxT = max_train;
    int devCount = 0;
    cudaError_t e = cudaGetDeviceCount(&devCount);
    if (e != cudaSuccess || device < 0 || device >= devCount) {
        cudaGetLastError();
        return mfail(m, ORBX_ERR_CUDA, e != cudaSuccess ? std::string("no CUDA device: ") + cudaGetErrorString(e) : "device ordinal out of range");
    }
    MCK(cudaSetDevice(device));
    cudaDeviceProp prop;
    MCK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return mfail(m, ORBX_ERR_CUDA, "liborbx is built for sm_100a only (no other code path exists)");
    m->smCount = prop.multiProcessorCount;
    MCK(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    MCK(cudaMalloc((void **)&m->dQ, (size_t)max_queries * 32));
    MCK(cudaMalloc((void **)&m->dT, (size_t)max_train * 32));
    MCK(cudaMalloc((void **)&m->dOut, (size_t)max_queries * sizeof(int4)));
    MCK(cudaMallocHost((void **)&m->hOut, (size_t)max_queries * sizeof(int4)));
    return ORBX_OK;
}

void orbm_destroy(orbm_matcher *m)
{
    if (!m) return;
    if (m->stream) { cudaSetDevice(m->device); cudaStreamSynchronize(m->stream); }
    if (m->dQ) cudaFree(m->dQ);
    if (m->dT) cudaFree(m->dT);
    if (m->dPart) cudaFree(m->dPart);
    if (m->dOut) cudaFree(m->dOut);
    if (m->hOut) cudaFreeHost(m->hOut);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
}

const char *orbm_last_error(const orbm_matcher *m) { return m ? m->err.c_str() : "null handle"; }

int orbm_set_train(orbm_matcher *m, const uint8_t *t, int nt)
{
    if (!m) return ORBX_ERR_ARG;
    if ((!t && nt > 0) || nt < 0 || nt > m->maxT) return mfail(m, ORBX_ERR_ARG, "bad train set");
    MCK(cudaSetDevice(m->device));
    if (nt > 0) MCK(cudaMemcpyAsync(m->dT, t, (size_t)nt * 32, cudaMemcpyHostToDevice, m->stream));
    MCK(cudaStreamSynchronize(m->stream));
    m->residentNt = nt;
    return ORBX_OK;
}

int orbm_knn2_resident(orbm_matcher *m, const uint8_t *q, int nq, int32_t *idx, int32_t *d1, int32_t *d2)
{
    if (!m) return ORBX_ERR_ARG;
    if (m->residentNt < 0) return mfail(m, ORBX_ERR_ARG, "no resident train set (call orbm_set_train)");
    if (!q || !idx || !d1 || !d2 || nq < 1 || nq > m->maxQ) return mfail(m, ORBX_ERR_ARG, "bad query block");
    MCK(cudaSetDevice(m->device));
    MCK(cudaMemcpyAsync(m->dQ, q, (size_t)nq * 32, cudaMemcpyHostToDevice, m->stream));
    int rc = enqueueKnn(m, m->dQ, nq, m->dT, m->residentNt, m->dOut, m->stream);
    if (rc != ORBX_OK) return rc;
    MCK(cudaMemcpyAsync(m->hOut, m->dOut, (size_t)nq * sizeof(int4), cudaMemcpyDeviceToHost, m->stream));
    MCK(cudaStreamSynchronize(m->stream));
    for (int i = 0; i < nq; i++) { idx[i] = m->hOut[i].x; d1[i] = m->hOut[i].y; d2[i] = m->hOut[i].z; }
    return ORBX_OK;
}

int orbm_knn2(orbm_matcher *m, const uint8_t *q, int nq, const uint8_t *t, int nt,
              int32_t *idx, int32_t *d1, int32_t *d2)
{
    int rc = orbm_set_train(m, t, nt);
    if (rc != ORBX_OK) return rc;
    return orbm_knn2_resident(m, q, nq, idx, d1, d2);
}

int orbm_knn2_device(orbm_matcher *m, const uint8_t *d_q, int nq, const uint8_t *d_t, int nt,
                     int32_t *d_out, void *stream)
{
    if (!m) return ORBX_ERR_ARG;
    if (!d_q || !d_out || (!d_t && nt > 0) || nq < 1 || nt < 0 || nt > (int)KNN_IDX_MASK - 1) return mfail(m, ORBX_ERR_ARG, "bad argument");
    MCK(cudaSetDevice(m->device));
    return enqueueKnn(m, d_q, nq, d_t, nt, (int4 *)d_out, stream ? (cudaStream_t)stream : m->stream);
}

int orbm_distance_pairs(orbm_matcher *m, const uint8_t *a, const uint8_t *b, int n, int32_t *out)
{
    if (!m) return ORBX_ERR_ARG;
    if (!a || !b || !out || n < 1 || n > m->maxQ || n > m->maxT) return mfail(m, ORBX_ERR_ARG, "bad argument");
    MCK(cudaSetDevice(m->device));
    MCK(cudaMemcpyAsync(m->dQ, a, (size_t)n * 32, cudaMemcpyHostToDevice, m->stream));
    MCK(cudaMemcpyAsync(m->dT, b, (size_t)n * 32, cudaMemcpyHostToDevice, m->stream));
    m->residentNt = -1;
    k_distance_pairs<<<(n + 127) / 128, 128, 0, m->stream>>>((const uint4 *)m->dQ, (const uint4 *)m->dT, n, (int *)m->dOut);
    MCK(cudaGetLastError());
    MCK(cudaMemcpyAsync(m->hOut, m->dOut, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    MCK(cudaStreamSynchronize(m->stream));
    memcpy(out, m->hOut, (size_t)n * sizeof(int));
    return ORBX_OK;
}

} // extern "C"
