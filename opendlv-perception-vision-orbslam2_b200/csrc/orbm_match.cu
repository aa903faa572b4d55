// orbm_match.cu -- Hamming kNN-2 matcher (include/orbx.h, orbm_* entry points).
//
// Reproduces ORBmatcher::DescriptorDistance (orbmatcher.cpp:1662-1677: 8 x 32-bit XOR + popcount)
// evaluated over every query x train pair, with the reference's best / second-best bookkeeping
// (orbmatcher.cpp:208-232: strict '<', start values 256 / -1; lowest index wins ties and the
// second best counts duplicates of the best).
//
// Kernel shape: INT pipe only (LOP3 carry-save adders + POPC + IADD3 + IMNMX) -- this is not a dense float
// contraction, tensor cores do not apply.  One query per thread held in 8 registers; the train
// set is streamed through shared memory in tiles that every thread of the CTA reads as 128-bit
// broadcasts; the grid is (query blocks) x (train chunks) so that 2000 queries still fill 148 SMs.
// The two smallest (distance << 22 | index) keys per query are kept -- keys are unique, so
// min-of-keys is "lowest index attains the minimum" and the second smallest key carries the
// second-best distance with multiplicity.  A merge kernel folds the per-chunk partials.
#include "../../include/orbx.h"

#include <cuda_runtime.h>

#include "orbx_smem_optin.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#define KNN_QB 128          // queries per CTA (one per thread)
#define KNN_TILE 256        // train descriptors per shared-memory tile (8 KB)
#define KNN_MIN_CHUNK 128   // shortest run of train descriptors worth a CTA of its own
#define KNN_IDX_BITS 22
#define KNN_IDX_MASK 0x3fffffu
#define KNN_INIT_KEY ((256u << KNN_IDX_BITS) | KNN_IDX_MASK)

// 256-bit Hamming distance.  POPC issues at a fraction of the LOP3 rate, so the eight XOR words are
// first compressed with carry-save adders (Harley-Seal): three words of equal weight become a sum word
// and a carry word of double weight (2 LOP3), leaving 4 POPC instead of 8.  The result is identical to
// the sum of the eight popcounts the reference computes (orbmatcher.cpp:1662-1677).
__device__ __forceinline__ unsigned csa_sum(unsigned a, unsigned b, unsigned c) { return a ^ b ^ c; }
__device__ __forceinline__ unsigned csa_carry(unsigned a, unsigned b, unsigned c) { return (a & b) | (c & (a | b)); }

__device__ __forceinline__ int hamming256(const uint4 &qa, const uint4 &qb, const uint4 &a, const uint4 &b)
{
    const unsigned x0 = qa.x ^ a.x, x1 = qa.y ^ a.y, x2 = qa.z ^ a.z, x3 = qa.w ^ a.w;
    const unsigned x4 = qb.x ^ b.x, x5 = qb.y ^ b.y, x6 = qb.z ^ b.z, x7 = qb.w ^ b.w;
    const unsigned s0 = csa_sum(x0, x1, x2), c0 = csa_carry(x0, x1, x2);
    const unsigned s1 = csa_sum(x3, x4, x5), c1 = csa_carry(x3, x4, x5);
    const unsigned s2 = csa_sum(s0, s1, x6), c2 = csa_carry(s0, s1, x6);
    const unsigned s3 = csa_sum(c0, c1, c2), c3 = csa_carry(c0, c1, c2);
    return __popc(s2) + __popc(x7) + 2 * __popc(s3) + 4 * __popc(c3);
}

// The two smallest keys (distance << 22 | train index) of one query over the train descriptors [t0, t1): the CTA's threads share
// the train tile in shared memory (every thread reads the same descriptor: a broadcast).  (Fetching the next tile into registers
// while the current one is scanned was measured and is no faster: four resident CTAs per SM cover each other's tile loads.)
__device__ __forceinline__ void knn_scan_chunk(const uint4 *__restrict__ t, int t0, int t1, const uint4 &qa, const uint4 &qb,
                                               uint4 *tile, int tid, unsigned &k1, unsigned &k2)
{
    for (int base = t0; base < t1; base += KNN_TILE) {
        const int cnt = min(KNN_TILE, t1 - base);
        __syncthreads();
        for (int i = tid; i < cnt * 2; i += KNN_QB) tile[i] = __ldg(&t[2 * (size_t)base + i]);
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < cnt; j++) {
            const uint4 a = tile[2 * j], b = tile[2 * j + 1];
            const int d = hamming256(qa, qb, a, b);
            const unsigned key = ((unsigned)d << KNN_IDX_BITS) | (unsigned)(base + j);
            k2 = min(k2, max(k1, key));
            k1 = min(k1, key);
        }
    }
}

__global__ void __launch_bounds__(KNN_QB, 4)
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
    knn_scan_chunk(t, t0, t1, qa, qb, tile, tid, k1, k2);
    if (qi < nq) part[(size_t)blockIdx.y * nq + qi] = make_uint2(k1, k2);
}

// Fold the per-chunk partial pairs of one query (part[c * stride], c < nchunks) into the two smallest keys.  KNN_FOLD loads are in
// flight per thread: with several hundred chunks (few queries, many chunks to fill the GPU) a load-by-load loop is a chain of
// L2 round trips -- 296 chunks x ~300 ns were most of the 30 us the sharded kernel spent behind its last chunk.  The last, partial batch
// loads from clamped positions instead of falling back to one load at a time.
#define KNN_FOLD 32
__device__ __forceinline__ void knn_fold_pair(unsigned &k1, unsigned &k2, unsigned x, unsigned y)
{
    // fold a sorted pair: the new pair is the two smallest of {k1, k2, x, y}
    k2 = min(k2, max(k1, x));
    k1 = min(k1, x);
    k2 = min(k2, max(k1, y));
    k1 = min(k1, y);
}
template <int B>
__device__ __forceinline__ void knn_fold(const uint2 *part, size_t stride, int nchunks, unsigned &k1, unsigned &k2)
{
    int c = 0;
    for (; c + B <= nchunks; c += B) {
        uint2 p[B];
#pragma unroll
        for (int u = 0; u < B; u++) p[u] = __ldcg(part + (size_t)(c + u) * stride);
#pragma unroll
        for (int u = 0; u < B; u++) knn_fold_pair(k1, k2, p[u].x, p[u].y);
    }
    if (c < nchunks) {
        // last batch: unconditional loads from clamped positions (so that they all go out together), the surplus replaced afterwards
        uint2 p[B];
#pragma unroll
        for (int u = 0; u < B; u++) p[u] = __ldcg(part + (size_t)min(c + u, nchunks - 1) * stride);
#pragma unroll
        for (int u = 0; u < B; u++) {
            const bool in = c + u < nchunks;
            knn_fold_pair(k1, k2, in ? p[u].x : KNN_INIT_KEY, in ? p[u].y : KNN_INIT_KEY);
        }
    }
}

// lanesPerQuery (1, 2, 4 or 8) adjacent lanes share a query: lane `sub` folds the chunks c = sub (mod lanesPerQuery), a butterfly
// joins the pairs.  With few queries and hundreds of chunks a thread per query leaves two CTAs walking a long chain of loads.
__global__ void __launch_bounds__(128, 1)
k_knn2_merge(const uint2 *__restrict__ part, int nq, int nchunks, int lanesPerQuery, int4 *__restrict__ out)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int L = lanesPerQuery, qi = g / L, sub = g & (L - 1);
    unsigned k1 = KNN_INIT_KEY, k2 = KNN_INIT_KEY;
    if (qi < nq) knn_fold<KNN_FOLD>(part + qi + (size_t)sub * nq, (size_t)nq * L, (nchunks - sub + L - 1) / L, k1, k2);
    for (int o = L >> 1; o > 0; o >>= 1) {
        const unsigned x = __shfl_xor_sync(0xffffffffu, k1, o), y = __shfl_xor_sync(0xffffffffu, k2, o);
        knn_fold_pair(k1, k2, x, y);
    }
    if (qi >= nq || sub) return;
    int d1 = (int)(k1 >> KNN_IDX_BITS), d2 = (int)(k2 >> KNN_IDX_BITS);
    int idx = (int)(k1 & KNN_IDX_MASK);
    if (d1 >= 256) { d1 = 256; idx = -1; }  // 'dist < bestDist1' with bestDist1 = 256 never fires
    if (d2 >= 256) d2 = 256;
    out[qi] = make_int4(idx, d1, d2, 0);
}

// ------------------------------------------------------------------------------------------
// Query-sharded kNN-2 with the result gather fused into the kernel (no collective library call).
// Rank r of n_ranks matches its block of queries against the whole (replicated) train set in ONE launch:
//   * the grid is the (query block x train chunk) grid of k_knn2_partial; a CTA writes its per-chunk partial pairs, and the
//     last CTA to finish a query block (ticket counter) folds the partials of that block -- there is no separate merge launch;
//   * the folding CTA stores the 16-byte records straight into the result window of EVERY rank (its own and the peers',
//     mapped through NVLink: peer access inside one process, CUDA IPC between processes) at the block's global offset;
//   * the CTA that finishes the rank's last query block publishes the rank's flag (= the call's epoch) in every window with
//     a system-scope release, then waits until all n_ranks flags of its OWN window carry the epoch.  When the launch completes
//     in stream order, this rank's window holds the records of all ranks.
// Windows carry two record buffers selected by the epoch's parity: a rank can be one call ahead of a peer that is still
// consuming the previous result, never two (it cannot pass the flag wait of call e before the peer has started call e).
// ------------------------------------------------------------------------------------------
#define KNN_MAX_RANKS 16
#define KNN_WIN_HDR 1024         // bytes in front of the records: one 32-bit flag per rank, 64 bytes apart
struct KnnPeers { unsigned char *win[KNN_MAX_RANKS]; };

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

#define KNN_GROUP 16             // chunks per fold group (first level of the in-kernel merge)

// what the CTA that finished a query block last does: fold, store everywhere, and -- for the rank's last block -- publish and wait.
// Kept out of line: inlined, its 32 loads in flight set the register allocation of the scan loop in front of it (3 % slower).
__device__ __noinline__ void knn_sharded_tail(const uint2 *folded, int nFolded, int nq, unsigned *counters, KnnPeers peers, int nRanks,
                                              int rank, int qOffset, int nqTotal, unsigned epoch, int *isLast)
{
    const int tid = threadIdx.x;
    const int qi = blockIdx.x * KNN_QB + tid;
    if (qi < nq) {
        unsigned k1 = KNN_INIT_KEY, k2 = KNN_INIT_KEY;
        knn_fold<KNN_FOLD>(folded + qi, (size_t)nq, nFolded, k1, k2);
        int d1 = (int)(k1 >> KNN_IDX_BITS), d2 = (int)(k2 >> KNN_IDX_BITS);
        int idx = (int)(k1 & KNN_IDX_MASK);
        if (d1 >= 256) { d1 = 256; idx = -1; }
        if (d2 >= 256) d2 = 256;
        const int4 rec = make_int4(idx, d1, d2, 0);
        const size_t off = KNN_WIN_HDR + ((size_t)(epoch & 1u) * nqTotal + (size_t)(qOffset + qi)) * sizeof(int4);
        for (int r = 0; r < nRanks; r++) *(int4 *)(peers.win[r] + off) = rec;     // own window and, through NVLink, every peer's
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        counters[blockIdx.x] = 0;                                                // ready for the next launch
        *isLast = atomicAdd(&counters[gridDim.x], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!*isLast) return;
    // ---- every query block of this rank is stored everywhere: publish, then wait for the other ranks
    if (tid == 0) counters[gridDim.x] = 0;
    if (tid < nRanks) st_release_sys((unsigned *)(peers.win[tid] + 64 * rank), epoch);
    if (tid < nRanks) {
        const unsigned *f = (const unsigned *)(peers.win[rank] + 64 * tid);
        unsigned long long t0, now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while ((int)(ld_acquire_sys(f) - epoch) < 0) {
            __nanosleep(64);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > 5000000000ull) {                                      // a peer never launched: give up after 5 s, loudly
                *(unsigned *)(peers.win[rank] + KNN_WIN_HDR - 4) = 1u + (unsigned)tid;
                break;
            }
        }
    }
}

// counters: [0, qBlocks) groups finished per query block, [qBlocks] query blocks finished, then qBlocks x nGroups chunk counts per
// group; all zero between launches.  gpart: nGroups x nq group partials.
__global__ void __launch_bounds__(KNN_QB, 4)
k_knn2_sharded(const uint4 *__restrict__ q, int nq, const uint4 *__restrict__ t, int nt, int chunk, uint2 *part, uint2 *gpart,
               unsigned *counters, KnnPeers peers, int nRanks, int rank, int qOffset, int nqTotal, unsigned epoch)
{
    __shared__ uint4 tile[KNN_TILE * 2];
    __shared__ int isLast;
    const int tid = threadIdx.x;
    const int qi = blockIdx.x * KNN_QB + tid;
    const int t0 = blockIdx.y * chunk;
    const int t1 = min(t0 + chunk, nt);
    uint4 qa = make_uint4(0, 0, 0, 0), qb = qa;
    if (qi < nq) { qa = __ldg(&q[2 * qi]); qb = __ldg(&q[2 * qi + 1]); }
    unsigned k1 = KNN_INIT_KEY, k2 = KNN_INIT_KEY;
    knn_scan_chunk(t, t0, t1, qa, qb, tile, tid, k1, k2);
    if (qi < nq) part[(size_t)blockIdx.y * nq + qi] = make_uint2(k1, k2);
    // ---- two-level fold of the per-chunk pairs, no separate launch and no serial walk over hundreds of chunks: the CTA that
    // finishes a group of KNN_GROUP chunks last folds that group (ticket pattern: the pairs are fenced before the ticket is
    // taken), the CTA that finishes a query block's last group folds the group results
    const int nchunks = (int)gridDim.y, nGroups = (nchunks + KNN_GROUP - 1) / KNN_GROUP;
    const int g = blockIdx.y / KNN_GROUP, gs = min(KNN_GROUP, nchunks - g * KNN_GROUP);
    unsigned *gcount = counters + gridDim.x + 1 + blockIdx.x * nGroups + g;
    __threadfence();
    __syncthreads();
    if (tid == 0) isLast = atomicAdd(gcount, 1u) == (unsigned)gs - 1u;
    __syncthreads();
    if (!isLast) return;
    __threadfence();
    if (nGroups > 1) {
        if (qi < nq) {
            k1 = KNN_INIT_KEY; k2 = KNN_INIT_KEY;
            knn_fold<KNN_GROUP>(part + (size_t)g * KNN_GROUP * nq + qi, (size_t)nq, gs, k1, k2);
            gpart[(size_t)g * nq + qi] = make_uint2(k1, k2);
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            *gcount = 0;                                                         // ready for the next launch
            isLast = atomicAdd(&counters[blockIdx.x], 1u) == (unsigned)nGroups - 1u;
        }
        __syncthreads();
        if (!isLast) return;
        __threadfence();
        knn_sharded_tail(gpart, nGroups, nq, counters, peers, nRanks, rank, qOffset, nqTotal, epoch, &isLast);
    } else {
        if (tid == 0) *gcount = 0;
        knn_sharded_tail(part, nchunks, nq, counters, peers, nRanks, rank, qOffset, nqTotal, epoch, &isLast);
    }
}

// a rank without queries of its own still takes part in the exchange: it publishes its flag and waits for the others
__global__ void k_knn2_publish(KnnPeers peers, int nRanks, int rank, unsigned epoch)
{
    const int tid = threadIdx.x;
    if (tid < nRanks) st_release_sys((unsigned *)(peers.win[tid] + 64 * rank), epoch);
    if (tid < nRanks) {
        const unsigned *f = (const unsigned *)(peers.win[rank] + 64 * tid);
        unsigned long long t0, now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while ((int)(ld_acquire_sys(f) - epoch) < 0) {
            __nanosleep(64);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > 5000000000ull) { *(unsigned *)(peers.win[rank] + KNN_WIN_HDR - 4) = 1u + (unsigned)tid; break; }
        }
    }
}

__global__ void k_distance_pairs(const uint4 *__restrict__ a, const uint4 *__restrict__ b, int n, int *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 a0 = a[2 * i], a1 = a[2 * i + 1], b0 = b[2 * i], b1 = b[2 * i + 1];
    out[i] = hamming256(a0, a1, b0, b1);
}

// Candidate-list matching (inner loop of SearchByProjection / SearchByBoW, orbmatcher.cpp:76-114):
// one warp per query, lanes stride over the query's CSR list.  Keys are (distance << 22 | position in
// the list): the reference's strict '<' updates keep the two smallest under exactly that order.
__global__ void __launch_bounds__(128)
k_knn2_csr(const uint4 *__restrict__ q, int nq, const uint4 *__restrict__ t, const int *__restrict__ offsets,
           const int *__restrict__ indices, int4 *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (qi >= nq) return;
    const uint4 qa = __ldg(&q[2 * qi]), qb = __ldg(&q[2 * qi + 1]);
    const int beg = offsets[qi], end = offsets[qi + 1];
    unsigned k1 = KNN_INIT_KEY, k2 = KNN_INIT_KEY;
    for (int p = beg + lane; p < end; p += 32) {
        const int ti = __ldg(&indices[p]);
        const uint4 a = __ldg(&t[2 * (size_t)ti]), b = __ldg(&t[2 * (size_t)ti + 1]);
        const int d = hamming256(qa, qb, a, b);
        const unsigned key = ((unsigned)d << KNN_IDX_BITS) | (unsigned)(p - beg);
        k2 = min(k2, max(k1, key));
        k1 = min(k1, key);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {   // merge two sorted pairs: the two smallest of four keys
        const unsigned o1 = __shfl_xor_sync(0xffffffffu, k1, o), o2 = __shfl_xor_sync(0xffffffffu, k2, o);
        k2 = min(min(k2, o2), max(k1, o1));
        k1 = min(k1, o1);
    }
    if (lane == 0) {
        int d1 = (int)(k1 >> KNN_IDX_BITS), d2 = (int)(k2 >> KNN_IDX_BITS);
        int i1 = d1 < 256 ? indices[beg + (int)(k1 & KNN_IDX_MASK)] : -1;
        int i2 = d2 < 256 ? indices[beg + (int)(k2 & KNN_IDX_MASK)] : -1;
        if (d1 >= 256) d1 = 256;
        if (d2 >= 256) d2 = 256;
        out[qi] = make_int4(i1, d1, d2, i2);
    }
}

// Distances of every (query, candidate) entry of a CSR list, in list order: what the sequential drivers
// (SearchByProjection & co., orbmatcher.cpp:76-114) need to replay their exclusion rules on the host without
// computing a single Hamming distance there.  One warp per query, lanes stride over its list.
__global__ void __launch_bounds__(128)
k_distance_csr(const uint4 *__restrict__ q, int nq, const uint4 *__restrict__ t, const int *__restrict__ offsets,
               const int *__restrict__ indices, int *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (qi >= nq) return;
    const uint4 qa = __ldg(&q[2 * qi]), qb = __ldg(&q[2 * qi + 1]);
    const int beg = offsets[qi], end = offsets[qi + 1];
    for (int p = beg + lane; p < end; p += 32) {
        const int ti = __ldg(&indices[p]);
        out[p] = hamming256(qa, qb, __ldg(&t[2 * (size_t)ti]), __ldg(&t[2 * (size_t)ti + 1]));
    }
}

// POPC issue-rate probe for the INT roofline of the matcher (SURVEY 8d asks for a measured R_popc):
// every thread runs 8 independent POPC->XOR chains; a CTA reports its own cycle count.
__global__ void __launch_bounds__(256)
k_popc_probe(unsigned seed, int iters, unsigned *sink, long long *cycles)
{
    unsigned a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = seed * (threadIdx.x + 1) + k * 0x9e3779b9u;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            unsigned p;
            asm volatile("popc.b32 %0, %1;" : "=r"(p) : "r"(a[k]));
            a[k] = p * 0x9e3779b1u + (unsigned)i;      // IMAD: the mixing runs on the FMA pipe, not beside POPC
        }
    }
    const long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= a[k];
    if (s == 0xdeadbeefu) sink[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// ------------------------------------------------------------------------------------------
// OrbMapPoint::ComputeDistinctiveDescriptors (orbmappoint.cpp:314-383): one warp per map point.  Lane = row i of the N x N distance matrix (rows strided by 32):
// it computes its row against every observed descriptor (the other descriptor is a broadcast load), keeps the row in
// shared memory (column-major over the lanes: conflict-free) and finds element (N-1)/2 of the sorted row by bisection
// on the value (distances lie in [0, 256]); the warp then takes the first row with the least median.
// ------------------------------------------------------------------------------------------
#define DD_WARPS 4
__global__ void __launch_bounds__(DD_WARPS * 32)
k_distinctive(const uint4 *__restrict__ desc, const int *__restrict__ offsets, const int *__restrict__ indices, int nPoints,
              int maxList, int2 *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * DD_WARPS + warp;
    if (p >= nPoints) return;
    int *ids = (int *)dsm + (size_t)warp * maxList;                                           // maxList per warp
    uint16_t *rows = (uint16_t *)((int *)dsm + (size_t)DD_WARPS * maxList) + (size_t)warp * maxList * 32;   // maxList x 32 per warp
    const int off = offsets[p], N = offsets[p + 1] - off;
    if (N <= 0) { if (lane == 0) out[p] = make_int2(-1, -1); return; }
    for (int j = lane; j < N; j += 32) ids[j] = indices[off + j];
    __syncwarp();
    const int k = (N - 1) >> 1;                      // (int)(0.5 * ((float)N - 1.0)), :367
    unsigned bestKey = 0xffffffffu;
    for (int i0 = 0; i0 < N; i0 += 32) {
        const int i = i0 + lane;
        if (i < N) {
            const uint4 a0 = __ldg(&desc[(size_t)ids[i] * 2]), a1 = __ldg(&desc[(size_t)ids[i] * 2 + 1]);
            for (int j = 0; j < N; j++) {
                const uint4 b0 = __ldg(&desc[(size_t)ids[j] * 2]), b1 = __ldg(&desc[(size_t)ids[j] * 2 + 1]);
                const int d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                              __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
                rows[j * 32 + lane] = (uint16_t)d;
            }
            int lo = 0, hi = 256;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                int cnt = 0;
                for (int j = 0; j < N; j++) cnt += rows[j * 32 + lane] <= mid;
                if (cnt >= k + 1) hi = mid; else lo = mid + 1;
            }
            bestKey = min(bestKey, (unsigned)lo << 16 | (unsigned)i);   // first row with the least median, :369-373
        }
    }
    bestKey = __reduce_min_sync(0xffffffffu, bestKey);
    if (lane == 0) out[p] = make_int2((int)(bestKey & 0xffffu), (int)(bestKey >> 16));
}

// ---------------------------------------------------------------------------------------------------------------------
// ORBmatcher::SearchByProjection(frame, map points, th) with the frame's grid (orbmatcher.cpp:42-124, orbframe.cpp:192-211,
// :308-393).  All float expressions are formed operation by operation (__fsub_rn / __fmul_rn: no FMA contraction), as the
// reference's x86-64 build evaluates them.
#define FG_COLS 64          // FRAME_GRID_COLS, orbframe.hpp:52
#define FG_ROWS 48          // FRAME_GRID_ROWS, orbframe.hpp:51
#define FG_CELLS (FG_COLS * FG_ROWS)
#define FG_SMEM_KEYS 8192

// One CTA: PosInGrid for every key point, cell histogram, exclusive scan, then warp 0 fills the cell lists in key-point
// order (AssignFeaturesToGrid pushes i = 0 .. N-1 in order, orbframe.cpp:202-209).  Cell index = ix * FG_ROWS + iy, the
// order GetFeaturesInArea walks the cells in (:339-341).
__global__ void __launch_bounds__(1024)
k_frame_grid(const orbx_keypoint *__restrict__ keys, int n, float minX, float minY, float invW, float invH,
             int *__restrict__ cellOf, int *__restrict__ cellStart, int *__restrict__ cellItems)
{
    __shared__ int cnt[FG_CELLS];
    __shared__ int wsum[32];
    __shared__ uint16_t scell[FG_SMEM_KEYS];          // cell of the first FG_SMEM_KEYS key points (0xffff = none): the fill loop's reads stay on chip
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int c = tid; c < FG_CELLS; c += 1024) cnt[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 1024) {
        const int px = (int)roundf(__fmul_rn(__fsub_rn(keys[i].x, minX), invW));      // :383
        const int py = (int)roundf(__fmul_rn(__fsub_rn(keys[i].y, minY), invH));      // :384
        const int c = (px < 0 || px >= FG_COLS || py < 0 || py >= FG_ROWS) ? -1 : px * FG_ROWS + py;   // :387
        if (i < FG_SMEM_KEYS) scell[i] = (uint16_t)c; else cellOf[i] = c;
        if (c >= 0) atomicAdd(&cnt[c], 1);
    }
    __syncthreads();
    const int a0 = cnt[3 * tid], a1 = cnt[3 * tid + 1], a2 = cnt[3 * tid + 2];
    const int s = a0 + a1 + a2;
    int inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = wsum[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += v; }
        wsum[lane] = wi - w;
    }
    __syncthreads();
    const int base = wsum[warp] + inc - s;
    cnt[3 * tid] = base; cnt[3 * tid + 1] = base + a0; cnt[3 * tid + 2] = base + a0 + a1;
    cellStart[3 * tid] = base; cellStart[3 * tid + 1] = base + a0; cellStart[3 * tid + 2] = base + a0 + a1;
    if (tid == 1023) cellStart[FG_CELLS] = base + s;
    __syncthreads();
    if (warp == 0) {
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            int c = -1;
            if (i < n) { if (i < FG_SMEM_KEYS) { const int v = scell[i]; c = v == 0xffff ? -1 : v; } else c = cellOf[i]; }
            const unsigned act = __ballot_sync(0xffffffffu, c >= 0);
            if (c >= 0) {
                const unsigned same = __match_any_sync(act, c);
                const int leader = __ffs(same) - 1;
                int b = 0;
                if (lane == leader) { b = cnt[c]; cnt[c] = b + __popc(same); }
                b = __shfl_sync(same, b, leader);
                cellItems[b + __popc(same & ((1u << lane) - 1u))] = i;
            }
            __syncwarp();
        }
    }
}

// OrbFrame::GetFeaturesInArea (orbframe.cpp:308-380) for one window: visit(idx, key point) is called for every feature the
// reference would push into `indices`, in its order (cell columns ix, cells iy inside a column, key-point order in a cell).
template <class Visit>
__device__ __forceinline__ void walk_area(const orbx_keypoint *__restrict__ keys, const int *__restrict__ cellStart,
                                          const int *__restrict__ cellItems, float minX, float minY, float invW, float invH,
                                          float x, float y, float r, int minLevel, int maxLevel, Visit visit)
{
    const int c0x = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, minX), r), invW)));             // :313
    if (c0x >= FG_COLS) return;
    const int c1x = min(FG_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, minX), r), invW)));    // :319
    if (c1x < 0) return;
    const int c0y = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, minY), r), invH)));             // :325
    if (c0y >= FG_ROWS) return;
    const int c1y = min(FG_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, minY), r), invH)));    // :331
    if (c1y < 0) return;
    const bool checkLevels = (minLevel > 0) || (maxLevel >= 0);                                          // :337
    for (int ix = c0x; ix <= c1x; ix++) {
        // the cells (ix, c0y .. c1y) are adjacent in the cell order: one contiguous run of the item list
        const int e0 = cellStart[ix * FG_ROWS + c0y], e1 = cellStart[ix * FG_ROWS + c1y + 1];
        for (int e = e0; e < e1; e++) {
            const int idx = cellItems[e];
            const orbx_keypoint kp = keys[idx];
            if (checkLevels) {
                if (kp.octave < minLevel) continue;                                    // :354
                if (maxLevel >= 0 && kp.octave > maxLevel) continue;                   // :358-363
            }
            const float dx = __fsub_rn(kp.x, x), dy = __fsub_rn(kp.y, y);
            if (!(fabsf(dx) < r && fabsf(dy) < r)) continue;                           // :370
            visit(idx, kp);
        }
    }
}

// One thread per map point: GetFeaturesInArea + the candidate loop + the acceptance, sequentially in the reference's order.
//
// The reference's loop over the map points is sequential (orbmatcher.cpp:48): an accepted map point is stored in
// F->m_mapPoints at once (:121), and when it has observations it hides its key point from every LATER map point of the same
// call (:87-89).  The device evaluates all map points at once and iterates to the fixpoint of that rule: in a round, map point
// i skips key point k when k was occupied on entry or when the previous round left an observed map point j < i on k
// (claimPrev[k] = least such j).  Map point 0 is final after round 1, map point i after round i + 1 at the latest, and a round
// that changes no choice has reproduced the sequential loop exactly; in practice two or three rounds.
__global__ void __launch_bounds__(128)
k_project_round(const orbx_keypoint *__restrict__ keys, const float *__restrict__ uRight, const uint8_t *__restrict__ occupied,
                const uint4 *__restrict__ desc, const int *__restrict__ cellStart, const int *__restrict__ cellItems,
                float minX, float minY, float invW, float invH, const uint4 *__restrict__ mpDesc, const float *__restrict__ mpX,
                const float *__restrict__ mpY, const int *__restrict__ mpLevel, const float *__restrict__ mpRadius,
                const uint8_t *__restrict__ mpObserved, int nMp, float nnRatio, int thHigh, const int *__restrict__ claimPrev,
                int *__restrict__ claimNext, int *__restrict__ mpMatch, int *__restrict__ changed)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nMp) return;
    const float x = mpX[i], y = mpY[i], r = mpRadius[i];
    const int level = mpLevel[i];
    const uint4 qa = mpDesc[2 * i], qb = mpDesc[2 * i + 1];
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;   // orbmatcher.cpp:75-79
    walk_area(keys, cellStart, cellItems, minX, minY, invW, invH, x, y, r, level - 1, level,   // :64-68
              [&](int idx, const orbx_keypoint &kp) {
        if (occupied && occupied[idx]) return;                                         // :87-89, state on entry
        if (claimPrev && claimPrev[idx] < i) return;                                   // :87-89, stored earlier in this call (:121)
        const float ur = uRight[idx];
        if (ur > 0.f && fabsf(__fsub_rn(x, ur)) > r) return;                           // :91-96
        const int dist = hamming256(qa, qb, desc[2 * idx], desc[2 * idx + 1]);
        if (dist < bestDist) {                                                         // :102-114
            bestDist2 = bestDist; bestDist = dist;
            bestLevel2 = bestLevel; bestLevel = kp.octave;
            bestIdx = idx;
        } else if (dist < bestDist2) {
            bestLevel2 = kp.octave; bestDist2 = dist;
        }
    });
    int choice = -1;
    if (bestDist <= thHigh &&                                                          // :116-123
        !(bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(nnRatio, (float)bestDist2)))
        choice = bestIdx;
    if (choice != mpMatch[i]) { mpMatch[i] = choice; *changed = 1; }
    if (choice >= 0 && claimNext && mpObserved[i]) atomicMin(&claimNext[choice], i);
}

// after the fixpoint: m_mapPoints[k] = the LAST map point accepted on k (the loop runs in map-point order, :121), nmatches (:122)
__global__ void __launch_bounds__(128)
k_project_finish(const int *__restrict__ mpMatch, int nMp, int *__restrict__ assigned, int *__restrict__ nMatches)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nMp) return;
    const int k = mpMatch[i];
    if (k < 0) return;
    atomicMax(&assigned[k], i);
    atomicAdd(nMatches, 1);
}

// GetFeaturesInArea for nq windows: pass 1 counts, one CTA scans the counts into CSR offsets, pass 2 writes the feature
// indices in the reference's order and, when query descriptors are given, DescriptorDistance of each.
__global__ void __launch_bounds__(128)
k_area_count(const orbx_keypoint *__restrict__ keys, const int *__restrict__ cellStart, const int *__restrict__ cellItems,
             float minX, float minY, float invW, float invH, const float *__restrict__ qX, const float *__restrict__ qY,
             const float *__restrict__ qR, const int *__restrict__ qMinL, const int *__restrict__ qMaxL, int nq, int *__restrict__ count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    int c = 0;
    walk_area(keys, cellStart, cellItems, minX, minY, invW, invH, qX[i], qY[i], qR[i], qMinL[i], qMaxL[i],
              [&](int, const orbx_keypoint &) { c++; });
    count[i] = c;
}

__global__ void __launch_bounds__(1024) k_area_scan(const int *__restrict__ count, int nq, int *__restrict__ offsets)
{
    __shared__ int wsum[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int i0 = 0; i0 < nq; i0 += 1024) {
        const int i = i0 + tid;
        const int c = i < nq ? count[i] : 0;
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = wsum[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += v; }
            wsum[lane] = wi - w;
        }
        __syncthreads();
        const int excl = carry + wsum[warp] + inc - c;
        if (i < nq) offsets[i] = excl;
        __syncthreads();
        if (tid == 1023) carry = excl + c;
        __syncthreads();
    }
    if (tid == 0) offsets[nq] = carry;
}

__global__ void __launch_bounds__(128)
k_area_fill(const orbx_keypoint *__restrict__ keys, const uint4 *__restrict__ desc, const int *__restrict__ cellStart,
            const int *__restrict__ cellItems, float minX, float minY, float invW, float invH, const uint4 *__restrict__ qDesc,
            const float *__restrict__ qX, const float *__restrict__ qY, const float *__restrict__ qR, const int *__restrict__ qMinL,
            const int *__restrict__ qMaxL, int nq, const int *__restrict__ offsets, int cap, int *__restrict__ indices,
            int *__restrict__ dist)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    if (offsets[nq] > cap) return;                       // the caller's arrays are too small: nothing is written
    int o = offsets[i];
    uint4 qa = make_uint4(0, 0, 0, 0), qb = qa;
    if (qDesc) { qa = qDesc[2 * i]; qb = qDesc[2 * i + 1]; }
    walk_area(keys, cellStart, cellItems, minX, minY, invW, invH, qX[i], qY[i], qR[i], qMinL[i], qMaxL[i],
              [&](int idx, const orbx_keypoint &) {
        indices[o] = idx;
        if (qDesc) dist[o] = hamming256(qa, qb, desc[2 * idx], desc[2 * idx + 1]);
        o++;
    });
}


struct orbm_matcher {
    int device = 0, maxQ = 0, maxT = 0, smCount = 148;
    cudaStream_t stream = nullptr;
    uint8_t *dQ = nullptr, *dT = nullptr;
    uint2 *dPart = nullptr; size_t partCap = 0;
    int *dCsr = nullptr; size_t csrCap = 0;
    int4 *dOut = nullptr;
    int4 *hOut = nullptr;
    int residentNt = -1;
    // orbm_distinctive: descriptor pool, CSR lists, results (grown on demand)
    uint8_t *ddDesc = nullptr; size_t ddCap = 0;
    int *ddCsr = nullptr; size_t ddCsrCap = 0;
    int2 *ddBest = nullptr, *ddHost = nullptr; int ddBestCap = 0;
    // orbm_search_by_projection: one workspace (frame + map points + grid + results), grown on demand
    uint8_t *spBuf = nullptr, *spHost = nullptr; size_t spCap = 0;
    // orbm_knn2_sharded: this rank's result window, the peers' mappings, ticket counters, call epoch
    unsigned char *win = nullptr; size_t winBytes = 0;
    int winRanks = 0, winRank = 0, winNq = 0;
    unsigned char *peerWin[16] = {};
    bool peerIpc[16] = {};
    unsigned *dCounters = nullptr; int countersCap = 0;
    unsigned epoch = 0;
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
    // four resident CTAs per SM; the chunk length is NOT rounded to whole tiles (the last tile of a chunk is simply shorter):
    // rounding 338 up to 512 left 392 CTAs for 148 SMs at 250 queries -- SMs with three CTAs next to SMs with two
    int want = std::max(1, (m->smCount * 4 + qb - 1) / qb);
    int maxChunks = std::max(1, (nt + KNN_MIN_CHUNK - 1) / KNN_MIN_CHUNK);
    int c = std::min(want, maxChunks);
    int per = std::max((nt + c - 1) / c, KNN_MIN_CHUNK);
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
    int lanes = 1;
    while (lanes < 8 && nq * lanes * 2 <= 8192 && chunks / (lanes * 2) >= 4) lanes *= 2;
    k_knn2_merge<<<(nq * lanes + 127) / 128, 128, 0, st>>>(m->dPart, nq, chunks, lanes, dout);
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
    if (m->dCsr) cudaFree(m->dCsr);
    if (m->dOut) cudaFree(m->dOut);
    if (m->hOut) cudaFreeHost(m->hOut);
    if (m->ddDesc) cudaFree(m->ddDesc);
    if (m->ddCsr) cudaFree(m->ddCsr);
    if (m->ddBest) cudaFree(m->ddBest);
    if (m->ddHost) cudaFreeHost(m->ddHost);
    if (m->spBuf) cudaFree(m->spBuf);
    if (m->spHost) cudaFreeHost(m->spHost);
    for (int r = 0; r < 16; r++) if (m->peerIpc[r] && m->peerWin[r]) cudaIpcCloseMemHandle(m->peerWin[r]);
    if (m->win) cudaFree(m->win);
    if (m->dCounters) cudaFree(m->dCounters);
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

int orbm_knn2_csr_device(orbm_matcher *m, const uint8_t *d_q, int nq, const uint8_t *d_t, const int32_t *d_offsets,
                         const int32_t *d_indices, int32_t *d_out, void *stream)
{
    if (!m) return ORBX_ERR_ARG;
    if (!d_q || !d_t || !d_offsets || !d_indices || !d_out || nq < 1) return mfail(m, ORBX_ERR_ARG, "bad argument");
    if (((uintptr_t)d_q | (uintptr_t)d_t | (uintptr_t)d_out) & 15) return mfail(m, ORBX_ERR_ARG, "device buffers must be 16-byte aligned");
    MCK(cudaSetDevice(m->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : m->stream;
    k_knn2_csr<<<(nq * 32 + 127) / 128, 128, 0, st>>>((const uint4 *)d_q, nq, (const uint4 *)d_t, d_offsets, d_indices, (int4 *)d_out);
    MCK(cudaGetLastError());
    return ORBX_OK;
}

int orbm_knn2_csr(orbm_matcher *m, const uint8_t *q, int nq, const uint8_t *t, int nt, const int32_t *offsets,
                  const int32_t *indices, int32_t *idx1, int32_t *d1, int32_t *idx2, int32_t *d2)
{
    if (!m) return ORBX_ERR_ARG;
    if (!q || !offsets || !indices || !idx1 || !d1 || !idx2 || !d2 || nq < 1 || nq > m->maxQ || nt < 0 || nt > m->maxT || (!t && nt > 0))
        return mfail(m, ORBX_ERR_ARG, "bad argument");
    const int nnz = offsets[nq];
    if (offsets[0] != 0 || nnz < 0) return mfail(m, ORBX_ERR_ARG, "offsets must start at 0 and be non-decreasing");
    for (int i = 0; i < nq; i++) {
        if (offsets[i + 1] < offsets[i]) return mfail(m, ORBX_ERR_ARG, "offsets must be non-decreasing");
        if (offsets[i + 1] - offsets[i] > (int)KNN_IDX_MASK) return mfail(m, ORBX_ERR_ARG, "candidate list too long");
    }
    for (int k = 0; k < nnz; k++) if (indices[k] < 0 || indices[k] >= nt) return mfail(m, ORBX_ERR_ARG, "candidate index out of range");
    MCK(cudaSetDevice(m->device));
    const size_t need = (size_t)nq + 1 + (size_t)nnz;
    if (need > m->csrCap) {
        if (m->dCsr) cudaFree(m->dCsr);
        m->dCsr = nullptr; m->csrCap = 0;
        MCK(cudaMalloc((void **)&m->dCsr, need * sizeof(int)));
        m->csrCap = need;
    }
    MCK(cudaMemcpyAsync(m->dQ, q, (size_t)nq * 32, cudaMemcpyHostToDevice, m->stream));
    if (nt > 0) MCK(cudaMemcpyAsync(m->dT, t, (size_t)nt * 32, cudaMemcpyHostToDevice, m->stream));
    m->residentNt = nt;
    MCK(cudaMemcpyAsync(m->dCsr, offsets, (size_t)(nq + 1) * sizeof(int), cudaMemcpyHostToDevice, m->stream));
    if (nnz > 0) MCK(cudaMemcpyAsync(m->dCsr + nq + 1, indices, (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice, m->stream));
    int rc = orbm_knn2_csr_device(m, m->dQ, nq, m->dT, m->dCsr, m->dCsr + nq + 1, (int32_t *)m->dOut, m->stream);
    if (rc != ORBX_OK) return rc;
    MCK(cudaMemcpyAsync(m->hOut, m->dOut, (size_t)nq * sizeof(int4), cudaMemcpyDeviceToHost, m->stream));
    MCK(cudaStreamSynchronize(m->stream));
    for (int i = 0; i < nq; i++) { idx1[i] = m->hOut[i].x; d1[i] = m->hOut[i].y; d2[i] = m->hOut[i].z; idx2[i] = m->hOut[i].w; }
    return ORBX_OK;
}

int orbm_distance_csr(orbm_matcher *m, const uint8_t *q, int nq, const uint8_t *t, int nt, const int32_t *offsets,
                      const int32_t *indices, int32_t *dist)
{
    if (!m) return ORBX_ERR_ARG;
    if (!q || !offsets || !indices || !dist || nq < 1 || nq > m->maxQ || nt < 0 || nt > m->maxT || (!t && nt > 0))
        return mfail(m, ORBX_ERR_ARG, "bad argument");
    const int nnz = offsets[nq];
    if (offsets[0] != 0 || nnz < 0) return mfail(m, ORBX_ERR_ARG, "offsets must start at 0 and be non-decreasing");
    for (int i = 0; i < nq; i++) if (offsets[i + 1] < offsets[i]) return mfail(m, ORBX_ERR_ARG, "offsets must be non-decreasing");
    for (int k = 0; k < nnz; k++) if (indices[k] < 0 || indices[k] >= nt) return mfail(m, ORBX_ERR_ARG, "candidate index out of range");
    if (nnz == 0) return ORBX_OK;
    MCK(cudaSetDevice(m->device));
    const size_t need = (size_t)nq + 1 + 2 * (size_t)nnz;          // offsets | indices | distances
    if (need > m->csrCap) {
        if (m->dCsr) cudaFree(m->dCsr);
        m->dCsr = nullptr; m->csrCap = 0;
        MCK(cudaMalloc((void **)&m->dCsr, need * sizeof(int)));
        m->csrCap = need;
    }
    int *dOff = m->dCsr, *dInd = m->dCsr + nq + 1, *dDist = dInd + nnz;
    MCK(cudaMemcpyAsync(m->dQ, q, (size_t)nq * 32, cudaMemcpyHostToDevice, m->stream));
    MCK(cudaMemcpyAsync(m->dT, t, (size_t)nt * 32, cudaMemcpyHostToDevice, m->stream));
    m->residentNt = nt;
    MCK(cudaMemcpyAsync(dOff, offsets, (size_t)(nq + 1) * sizeof(int), cudaMemcpyHostToDevice, m->stream));
    MCK(cudaMemcpyAsync(dInd, indices, (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice, m->stream));
    k_distance_csr<<<(nq * 32 + 127) / 128, 128, 0, m->stream>>>((const uint4 *)m->dQ, nq, (const uint4 *)m->dT, dOff, dInd, dDist);
    MCK(cudaGetLastError());
    MCK(cudaMemcpyAsync(dist, dDist, (size_t)nnz * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    MCK(cudaStreamSynchronize(m->stream));
    return ORBX_OK;
}

int orbm_distinctive(orbm_matcher *m, const uint8_t *desc, int n_desc, const int32_t *offsets, const int32_t *indices, int n_points,
                     int32_t *best, int32_t *median)
{
    if (!m) return ORBX_ERR_ARG;
    if (!desc || !offsets || !indices || !best || n_desc < 1 || n_points < 1) return mfail(m, ORBX_ERR_ARG, "bad argument");
    if (offsets[0] != 0) return mfail(m, ORBX_ERR_ARG, "offsets must start at 0");
    int maxList = 1;
    for (int p = 0; p < n_points; p++) {
        if (offsets[p + 1] < offsets[p]) return mfail(m, ORBX_ERR_ARG, "offsets must be non-decreasing");
        maxList = std::max(maxList, offsets[p + 1] - offsets[p]);
    }
    if (maxList > 768) return mfail(m, ORBX_ERR_ARG, "more than 768 observations of one map point");
    const int nnz = offsets[n_points];
    for (int k = 0; k < nnz; k++) if (indices[k] < 0 || indices[k] >= n_desc) return mfail(m, ORBX_ERR_ARG, "descriptor index out of range");
    MCK(cudaSetDevice(m->device));
    if ((size_t)n_desc * 32 > m->ddCap) {
        cudaFree(m->ddDesc); m->ddDesc = nullptr; m->ddCap = 0;
        MCK(cudaMalloc((void **)&m->ddDesc, (size_t)n_desc * 32));
        m->ddCap = (size_t)n_desc * 32;
    }
    const size_t need = (size_t)n_points + 1 + (size_t)std::max(nnz, 1);
    if (need > m->ddCsrCap) {
        cudaFree(m->ddCsr); m->ddCsr = nullptr; m->ddCsrCap = 0;
        MCK(cudaMalloc((void **)&m->ddCsr, need * sizeof(int)));
        m->ddCsrCap = need;
    }
    if (n_points > m->ddBestCap) {
        cudaFree(m->ddBest); if (m->ddHost) cudaFreeHost(m->ddHost);
        m->ddBest = nullptr; m->ddHost = nullptr; m->ddBestCap = 0;
        MCK(cudaMalloc((void **)&m->ddBest, (size_t)n_points * sizeof(int2)));
        MCK(cudaMallocHost((void **)&m->ddHost, (size_t)n_points * sizeof(int2)));
        m->ddBestCap = n_points;
    }
    MCK(cudaMemcpyAsync(m->ddDesc, desc, (size_t)n_desc * 32, cudaMemcpyHostToDevice, m->stream));
    MCK(cudaMemcpyAsync(m->ddCsr, offsets, (size_t)(n_points + 1) * sizeof(int), cudaMemcpyHostToDevice, m->stream));
    if (nnz > 0) MCK(cudaMemcpyAsync(m->ddCsr + n_points + 1, indices, (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice, m->stream));
    const size_t smem = (size_t)DD_WARPS * maxList * (sizeof(int) + 32 * sizeof(uint16_t));
    MCK(orbx_raise_dyn_smem((const void *)k_distinctive, smem));
    k_distinctive<<<(n_points + DD_WARPS - 1) / DD_WARPS, DD_WARPS * 32, smem, m->stream>>>((const uint4 *)m->ddDesc, m->ddCsr, m->ddCsr + n_points + 1,
                                                                                              n_points, maxList, m->ddBest);
    MCK(cudaGetLastError());
    MCK(cudaMemcpyAsync(m->ddHost, m->ddBest, (size_t)n_points * sizeof(int2), cudaMemcpyDeviceToHost, m->stream));
    MCK(cudaStreamSynchronize(m->stream));
    for (int p = 0; p < n_points; p++) { best[p] = m->ddHost[p].x; if (median) median[p] = m->ddHost[p].y; }
    return ORBX_OK;
}


// ---- query-sharded kNN-2 with the fused gather (k_knn2_sharded)
int orbm_window_create(orbm_matcher *m, int nq_total, int n_ranks, int rank, void *ipc_handle)
{
    if (!m) return ORBX_ERR_ARG;
    if (nq_total < 1 || n_ranks < 1 || n_ranks > KNN_MAX_RANKS || rank < 0 || rank >= n_ranks) return mfail(m, ORBX_ERR_ARG, "bad argument");
    MCK(cudaSetDevice(m->device));
    for (int r = 0; r < 16; r++) {
        if (m->peerIpc[r] && m->peerWin[r]) cudaIpcCloseMemHandle(m->peerWin[r]);
        m->peerWin[r] = nullptr; m->peerIpc[r] = false;
    }
    if (m->win) { cudaFree(m->win); m->win = nullptr; }
    m->winBytes = KNN_WIN_HDR + (size_t)2 * nq_total * sizeof(int4);
    MCK(cudaMalloc((void **)&m->win, m->winBytes));
    MCK(cudaMemset(m->win, 0, m->winBytes));
    m->winRanks = n_ranks; m->winRank = rank; m->winNq = nq_total; m->epoch = 0;
    m->peerWin[rank] = m->win;
    // scratch of k_knn2_sharded for the largest call this window admits, allocated HERE: cudaFree / cudaMalloc between two
    // ranks' launches would wait for a kernel that is itself waiting for the other rank's flag (two slots on one device)
    {
        // likewise the two kernels are loaded now: with lazy module loading the FIRST launch of a kernel may have to wait for the
        // device to drain, i.e. for a peer slot's kernel that is waiting for this one
        cudaFuncAttributes fa;
        MCK(cudaFuncGetAttributes(&fa, (const void *)k_knn2_sharded));
        MCK(cudaFuncGetAttributes(&fa, (const void *)k_knn2_publish));
        const int qbMax = (nq_total + KNN_QB - 1) / KNN_QB;
        const size_t needPart = (size_t)m->smCount * 4 * KNN_QB + (size_t)(qbMax + 1) * KNN_QB;     // >= chunks * nq_local of knnGrid
        const size_t need = needPart + needPart / KNN_GROUP + (size_t)(qbMax + 1) * KNN_QB;            // + the group partials behind them
        if (need > m->partCap) {
            if (m->dPart) cudaFree(m->dPart);
            m->dPart = nullptr; m->partCap = 0;
            MCK(cudaMalloc((void **)&m->dPart, need * sizeof(uint2)));
            m->partCap = need;
        }
        // per query block one counter and one per group of chunks (chunks <= smCount * 4 / qb + 1), one for the rank
        const int nCounters = qbMax + 1 + (m->smCount * 4 + qbMax) / KNN_GROUP + 2 * qbMax + 16;
        if (nCounters > m->countersCap) {
            if (m->dCounters) cudaFree(m->dCounters);
            m->dCounters = nullptr; m->countersCap = 0;
            MCK(cudaMalloc((void **)&m->dCounters, (size_t)nCounters * sizeof(unsigned)));
            MCK(cudaMemset(m->dCounters, 0, (size_t)nCounters * sizeof(unsigned)));
            m->countersCap = nCounters;
        }
        // one empty launch of each (no ranks, no queries) so that nothing is left to load at the first real call
        KnnPeers none;
        for (int r = 0; r < KNN_MAX_RANKS; r++) none.win[r] = nullptr;
        k_knn2_publish<<<1, 32, 0, m->stream>>>(none, 0, 0, 0u);
        k_knn2_sharded<<<dim3(1, 1), KNN_QB, 0, m->stream>>>(nullptr, 0, nullptr, 0, KNN_TILE, m->dPart, m->dPart, m->dCounters, none, 0, 0, 0, nq_total, 0u);
        MCK(cudaGetLastError());
        MCK(cudaStreamSynchronize(m->stream));
    }
    if (ipc_handle) {
        cudaIpcMemHandle_t hd;
        MCK(cudaIpcGetMemHandle(&hd, m->win));
        static_assert(sizeof(hd) == 64, "CUDA IPC handles are 64 bytes");
        memcpy(ipc_handle, &hd, sizeof hd);
    }
    return ORBX_OK;
}

int orbm_window_attach_ipc(orbm_matcher *m, int peer_rank, const void *ipc_handle)
{
    if (!m) return ORBX_ERR_ARG;
    if (!m->win || !ipc_handle || peer_rank < 0 || peer_rank >= m->winRanks || peer_rank == m->winRank)
        return mfail(m, ORBX_ERR_ARG, "no window, or bad peer rank");
    MCK(cudaSetDevice(m->device));
    cudaIpcMemHandle_t hd;
    memcpy(&hd, ipc_handle, sizeof hd);
    void *p = nullptr;
    MCK(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    m->peerWin[peer_rank] = (unsigned char *)p; m->peerIpc[peer_rank] = true;
    return ORBX_OK;
}

int orbm_window_attach_peer(orbm_matcher *m, int peer_rank, orbm_matcher *peer)
{
    if (!m) return ORBX_ERR_ARG;
    if (!m->win || !peer || !peer->win || peer_rank < 0 || peer_rank >= m->winRanks || peer_rank == m->winRank ||
        peer->winRank != peer_rank || peer->winNq != m->winNq || peer->winRanks != m->winRanks)
        return mfail(m, ORBX_ERR_ARG, "windows of the two matchers do not belong to one group");
    MCK(cudaSetDevice(m->device));
    if (peer->device != m->device) {
        int can = 0;
        MCK(cudaDeviceCanAccessPeer(&can, m->device, peer->device));
        if (!can) return mfail(m, ORBX_ERR_CUDA, "no peer access between the two devices");
        cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return mfail(m, ORBX_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        cudaGetLastError();
    }
    m->peerWin[peer_rank] = peer->win; m->peerIpc[peer_rank] = false;
    return ORBX_OK;
}

int orbm_knn2_sharded(orbm_matcher *m, const uint8_t *d_q, int nq_local, int q_offset, const uint8_t *d_t, int nt, void *stream)
{
    if (!m) return ORBX_ERR_ARG;
    if (!m->win) return mfail(m, ORBX_ERR_ARG, "no result window (orbm_window_create)");
    if ((nq_local > 0 && (!d_q || !d_t)) || nq_local < 0 || q_offset < 0 || q_offset + nq_local > m->winNq || nt < 1 || nt > (int)KNN_IDX_MASK - 1)
        return mfail(m, ORBX_ERR_ARG, "bad argument");
    if (((uintptr_t)d_q | (uintptr_t)d_t) & 15) return mfail(m, ORBX_ERR_ARG, "device buffers must be 16-byte aligned");
    for (int r = 0; r < m->winRanks; r++) if (!m->peerWin[r]) return mfail(m, ORBX_ERR_ARG, "a peer's window has not been attached");
    MCK(cudaSetDevice(m->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : m->stream;
    if (nq_local == 0) {          // no queries of its own (fewer queries than ranks): the rank still publishes and waits
        KnnPeers pe;
        for (int r = 0; r < KNN_MAX_RANKS; r++) pe.win[r] = r < m->winRanks ? m->peerWin[r] : nullptr;
        m->epoch++;
        k_knn2_publish<<<1, 32, 0, st>>>(pe, m->winRanks, m->winRank, m->epoch);
        MCK(cudaGetLastError());
        return ORBX_OK;
    }
    int qb, chunks, chunk;
    knnGrid(m, nq_local, nt, &qb, &chunks, &chunk);
    const int nGroups = (chunks + KNN_GROUP - 1) / KNN_GROUP;
    if ((size_t)(chunks + nGroups) * nq_local > m->partCap || qb + 1 + qb * nGroups > m->countersCap)
        return mfail(m, ORBX_ERR_ARG, "scratch of the window is too small for this call");
    KnnPeers peers;
    for (int r = 0; r < KNN_MAX_RANKS; r++) peers.win[r] = r < m->winRanks ? m->peerWin[r] : nullptr;
    m->epoch++;
    k_knn2_sharded<<<dim3(qb, chunks), KNN_QB, 0, st>>>((const uint4 *)d_q, nq_local, (const uint4 *)d_t, nt, chunk, m->dPart,
                                                        m->dPart + (size_t)chunks * nq_local, m->dCounters,
                                                        peers, m->winRanks, m->winRank, q_offset, m->winNq, m->epoch);
    MCK(cudaGetLastError());
    return ORBX_OK;
}

int orbm_window_status(orbm_matcher *m, void *stream)
{
    if (!m) return ORBX_ERR_ARG;
    if (!m->win) return mfail(m, ORBX_ERR_ARG, "no window");
    MCK(cudaSetDevice(m->device));
    MCK(cudaStreamSynchronize(stream ? (cudaStream_t)stream : m->stream));
    unsigned bad = 0;
    MCK(cudaMemcpy(&bad, m->win + KNN_WIN_HDR - 4, 4, cudaMemcpyDeviceToHost));
    if (bad) {
        char msg[96];
        snprintf(msg, sizeof msg, "orbm_knn2_sharded: rank %u did not deliver its records within 5 s", bad - 1);
        return mfail(m, ORBX_ERR_CUDA, msg);
    }
    return ORBX_OK;
}

int orbm_window_fetch(orbm_matcher *m, void *stream, int32_t *records, int nq)
{
    if (!m) return ORBX_ERR_ARG;
    if (!m->win || !records || nq < 1 || nq > m->winNq || m->epoch == 0) return mfail(m, ORBX_ERR_ARG, "no window, no call yet or bad count");
    int rc = orbm_window_status(m, stream);
    if (rc != ORBX_OK) return rc;
    MCK(cudaMemcpy(records, m->win + KNN_WIN_HDR + (size_t)(m->epoch & 1u) * m->winNq * sizeof(int4), (size_t)nq * sizeof(int4), cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

int orbm_window_records(orbm_matcher *m, const int32_t **d_records)
{
    if (!m) return ORBX_ERR_ARG;
    if (!m->win || !d_records || m->epoch == 0) return mfail(m, ORBX_ERR_ARG, "no window or no call yet");
    *d_records = (const int32_t *)(m->win + KNN_WIN_HDR + (size_t)(m->epoch & 1u) * m->winNq * sizeof(int4));
    return ORBX_OK;
}


int orbm_measure_popc(orbm_matcher *m, double *popc_per_clk_per_sm)
{
    if (!m || !popc_per_clk_per_sm) return ORBX_ERR_ARG;
    MCK(cudaSetDevice(m->device));
    const int perSm = 8, blocks = m->smCount * perSm, iters = 4096;
    long long *dCyc = nullptr; unsigned *dSink = nullptr;
    MCK(cudaMalloc((void **)&dCyc, sizeof(long long) * blocks));
    MCK(cudaMalloc((void **)&dSink, sizeof(unsigned)));
    for (int rep = 0; rep < 2; rep++) k_popc_probe<<<blocks, 256, 0, m->stream>>>(12345u + rep, iters, dSink, dCyc);
    MCK(cudaGetLastError());
    std::vector<long long> cyc(blocks);
    MCK(cudaMemcpyAsync(cyc.data(), dCyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost, m->stream));
    MCK(cudaStreamSynchronize(m->stream));
    cudaFree(dCyc); cudaFree(dSink);
    double mean = 0;
    for (long long c : cyc) mean += (double)c;
    mean /= blocks;
    // perSm CTAs of 256 threads are co-resident on every SM for the whole probe
    *popc_per_clk_per_sm = (double)perSm * 256.0 * iters * 8.0 / mean;
    return ORBX_OK;
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


int orbm_search_by_projection(orbm_matcher *m, const orbm_frame_view *frame, const uint8_t *mp_desc, const float *mp_x,
                              const float *mp_y, const int32_t *mp_level, const float *mp_radius, const uint8_t *mp_observed,
                              int n_mp, float nnratio, int th_high, int32_t *mp_match, int32_t *assigned, int32_t *nmatches)
{
    if (!m) return ORBX_ERR_ARG;
    if (!frame || !mp_desc || !mp_x || !mp_y || !mp_level || !mp_radius || !mp_match || !assigned || !nmatches || n_mp < 1 ||
        frame->n < 0 || (frame->n > 0 && (!frame->keys || !frame->u_right || !frame->desc)) || th_high < 0 || th_high > 255 ||
        !(frame->max_x > frame->min_x) || !(frame->max_y > frame->min_y))
        return mfail(m, ORBX_ERR_ARG, "bad argument");
    const int n = frame->n;
    *nmatches = 0;
    for (int i = 0; i < n_mp; i++) mp_match[i] = -1;
    for (int k = 0; k < n; k++) assigned[k] = -1;
    if (n == 0) return ORBX_OK;
    MCK(cudaSetDevice(m->device));
    // m_gridElementWidthInverse / HeightInverse, orbframe.cpp:179-180
    const float invW = (float)FG_COLS / (frame->max_x - frame->min_x), invH = (float)FG_ROWS / (frame->max_y - frame->min_y);
    // workspace layout: [inputs, copied with ONE H2D from a pinned mirror] [device-only grid] [results, ONE D2H]
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t o = 0;
    const size_t oDesc = o;    o += al((size_t)n * 32);
    const size_t oMpDesc = o;  o += al((size_t)n_mp * 32);
    const size_t oKeys = o;    o += al((size_t)n * sizeof(orbx_keypoint));
    const size_t oUr = o;      o += al((size_t)n * 4);
    const size_t oOcc = o;     o += al((size_t)n);
    const size_t oMpX = o;     o += al((size_t)n_mp * 4);
    const size_t oMpY = o;     o += al((size_t)n_mp * 4);
    const size_t oMpR = o;     o += al((size_t)n_mp * 4);
    const size_t oMpL = o;     o += al((size_t)n_mp * 4);
    const size_t oMpObs = o;   o += al((size_t)n_mp);
    const size_t inBytes = o;
    const size_t oCellOf = o;  o += al((size_t)n * 4);
    const size_t oItems = o;   o += al((size_t)n * 4);
    const size_t oStart = o;   o += al((size_t)(FG_CELLS + 1) * 4);
    const size_t oClaimA = o;  o += al((size_t)n * 4);             // least observed map point stored on a key point, two rounds
    const size_t oClaimB = o;  o += al((size_t)n * 4);
    const size_t oRes = o;
    const size_t oMatch = o;   o += al((size_t)n_mp * 4);
    const size_t oAsg = o;     o += al((size_t)n * 4 + 8);          // assigned[n] | nmatches | changed
    const size_t resBytes = o - oRes;
    if (o > m->spCap) {
        if (m->spBuf) cudaFree(m->spBuf);
        if (m->spHost) cudaFreeHost(m->spHost);
        m->spBuf = nullptr; m->spHost = nullptr; m->spCap = 0;
        MCK(cudaMalloc((void **)&m->spBuf, o));
        MCK(cudaMallocHost((void **)&m->spHost, o));
        m->spCap = o;
    }
    uint8_t *b = m->spBuf, *hb = m->spHost;
    cudaStream_t st = m->stream;
    memcpy(hb + oDesc, frame->desc, (size_t)n * 32);
    memcpy(hb + oMpDesc, mp_desc, (size_t)n_mp * 32);
    memcpy(hb + oKeys, frame->keys, (size_t)n * sizeof(orbx_keypoint));
    memcpy(hb + oUr, frame->u_right, (size_t)n * 4);
    if (frame->occupied) memcpy(hb + oOcc, frame->occupied, (size_t)n);
    memcpy(hb + oMpX, mp_x, (size_t)n_mp * 4);
    memcpy(hb + oMpY, mp_y, (size_t)n_mp * 4);
    memcpy(hb + oMpR, mp_radius, (size_t)n_mp * 4);
    memcpy(hb + oMpL, mp_level, (size_t)n_mp * 4);
    bool anyObserved = false;
    if (mp_observed) {
        memcpy(hb + oMpObs, mp_observed, (size_t)n_mp);
        for (int i = 0; i < n_mp && !anyObserved; i++) anyObserved = mp_observed[i] != 0;
    }
    MCK(cudaMemcpyAsync(b, hb, inBytes, cudaMemcpyHostToDevice, st));
    MCK(cudaMemsetAsync(b + oMatch, 0xFF, (size_t)n_mp * 4, st));                 // choices start at -1
    MCK(cudaMemsetAsync(b + oAsg, 0xFF, (size_t)n * 4, st));
    MCK(cudaMemsetAsync(b + oAsg + (size_t)n * 4, 0, 8, st));
    k_frame_grid<<<1, 1024, 0, st>>>((const orbx_keypoint *)(b + oKeys), n, frame->min_x, frame->min_y, invW, invH,
                                     (int *)(b + oCellOf), (int *)(b + oStart), (int *)(b + oItems));
    MCK(cudaGetLastError());
    int *dChanged = (int *)(b + oAsg + (size_t)n * 4 + 4);
    auto round = [&](const int *claimPrev, int *claimNext) -> cudaError_t {
        k_project_round<<<(n_mp + 127) / 128, 128, 0, st>>>(
            (const orbx_keypoint *)(b + oKeys), (const float *)(b + oUr), frame->occupied ? b + oOcc : nullptr, (const uint4 *)(b + oDesc),
            (const int *)(b + oStart), (const int *)(b + oItems), frame->min_x, frame->min_y, invW, invH, (const uint4 *)(b + oMpDesc),
            (const float *)(b + oMpX), (const float *)(b + oMpY), (const int *)(b + oMpL), (const float *)(b + oMpR), b + oMpObs, n_mp,
            nnratio, th_high, claimPrev, claimNext, (int *)(b + oMatch), dChanged);
        return cudaGetLastError();
    };
    if (!anyObserved) {
        MCK(round(nullptr, nullptr));             // no accepted map point can hide a key point: one round is the whole loop
    } else {
        // rounds in groups of three, the "changed" flag of a group's last round read back; a round past the fixpoint changes nothing
        int *claim[2] = {(int *)(b + oClaimA), (int *)(b + oClaimB)};
        MCK(cudaMemsetAsync(claim[0], 0x7F, (size_t)n * 4, st));                   // 0x7f7f7f7f: no claimant
        int cur = 0, rounds = 0;
        for (;;) {
            for (int g = 0; g < 3; g++) {
                MCK(cudaMemsetAsync(claim[cur ^ 1], 0x7F, (size_t)n * 4, st));
                if (g == 2) MCK(cudaMemsetAsync(dChanged, 0, 4, st));
                MCK(round(claim[cur], claim[cur ^ 1]));
                cur ^= 1; rounds++;
            }
            int changedHost = 0;
            MCK(cudaMemcpyAsync(&changedHost, dChanged, 4, cudaMemcpyDeviceToHost, st));
            MCK(cudaStreamSynchronize(st));
            if (!changedHost) break;
            if (rounds > n_mp + 3) return mfail(m, ORBX_ERR_CUDA, "projection search did not reach its fixpoint");
        }
    }
    k_project_finish<<<(n_mp + 127) / 128, 128, 0, st>>>((const int *)(b + oMatch), n_mp, (int *)(b + oAsg), (int *)(b + oAsg + (size_t)n * 4));
    MCK(cudaGetLastError());
    MCK(cudaMemcpyAsync(hb + oRes, b + oRes, resBytes, cudaMemcpyDeviceToHost, st));
    MCK(cudaStreamSynchronize(st));
    memcpy(mp_match, hb + oMatch, (size_t)n_mp * 4);
    memcpy(assigned, hb + oAsg, (size_t)n * 4);
    memcpy(nmatches, hb + oAsg + (size_t)n * 4, 4);
    return ORBX_OK;
}


int orbm_area_distances(orbm_matcher *m, const orbm_frame_view *frame, const uint8_t *q_desc, const float *q_x, const float *q_y,
                        const float *q_r, const int32_t *q_min_level, const int32_t *q_max_level, int nq, int32_t *offsets,
                        int32_t *indices, int32_t *dist, int cap, int32_t *n_entries)
{
    if (!m) return ORBX_ERR_ARG;
    if (!frame || !q_x || !q_y || !q_r || !q_min_level || !q_max_level || !offsets || !n_entries || nq < 1 || cap < 0 ||
        (cap > 0 && (!indices || (q_desc && !dist))) || frame->n < 0 || (frame->n > 0 && (!frame->keys || !frame->desc)) ||
        !(frame->max_x > frame->min_x) || !(frame->max_y > frame->min_y))
        return mfail(m, ORBX_ERR_ARG, "bad argument");
    const int n = frame->n;
    *n_entries = 0;
    for (int i = 0; i <= nq; i++) offsets[i] = 0;
    if (n == 0) return ORBX_OK;
    MCK(cudaSetDevice(m->device));
    const float invW = (float)FG_COLS / (frame->max_x - frame->min_x), invH = (float)FG_ROWS / (frame->max_y - frame->min_y);
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t o = 0;
    const size_t oDesc = o;    o += al((size_t)n * 32);
    const size_t oQDesc = o;   o += al((size_t)nq * 32);
    const size_t oKeys = o;    o += al((size_t)n * sizeof(orbx_keypoint));
    const size_t oQX = o;      o += al((size_t)nq * 4);
    const size_t oQY = o;      o += al((size_t)nq * 4);
    const size_t oQR = o;      o += al((size_t)nq * 4);
    const size_t oQL0 = o;     o += al((size_t)nq * 4);
    const size_t oQL1 = o;     o += al((size_t)nq * 4);
    const size_t inBytes = o;
    const size_t oCellOf = o;  o += al((size_t)n * 4);
    const size_t oItems = o;   o += al((size_t)n * 4);
    const size_t oStart = o;   o += al((size_t)(FG_CELLS + 1) * 4);
    const size_t oCount = o;   o += al((size_t)nq * 4);
    const size_t oRes = o;
    const size_t oOff = o;     o += al((size_t)(nq + 1) * 4);
    const size_t oInd = o;     o += al((size_t)cap * 4);
    const size_t oDist = o;    o += al((size_t)cap * 4);
    if (o > m->spCap) {
        if (m->spBuf) cudaFree(m->spBuf);
        if (m->spHost) cudaFreeHost(m->spHost);
        m->spBuf = nullptr; m->spHost = nullptr; m->spCap = 0;
        MCK(cudaMalloc((void **)&m->spBuf, o));
        MCK(cudaMallocHost((void **)&m->spHost, o));
        m->spCap = o;
    }
    uint8_t *b = m->spBuf, *hb = m->spHost;
    cudaStream_t st = m->stream;
    memcpy(hb + oDesc, frame->desc, (size_t)n * 32);
    if (q_desc) memcpy(hb + oQDesc, q_desc, (size_t)nq * 32);
    memcpy(hb + oKeys, frame->keys, (size_t)n * sizeof(orbx_keypoint));
    memcpy(hb + oQX, q_x, (size_t)nq * 4);
    memcpy(hb + oQY, q_y, (size_t)nq * 4);
    memcpy(hb + oQR, q_r, (size_t)nq * 4);
    memcpy(hb + oQL0, q_min_level, (size_t)nq * 4);
    memcpy(hb + oQL1, q_max_level, (size_t)nq * 4);
    MCK(cudaMemcpyAsync(b, hb, inBytes, cudaMemcpyHostToDevice, st));
    const orbx_keypoint *dKeys = (const orbx_keypoint *)(b + oKeys);
    const int *dStart = (const int *)(b + oStart), *dItems = (const int *)(b + oItems);
    k_frame_grid<<<1, 1024, 0, st>>>(dKeys, n, frame->min_x, frame->min_y, invW, invH, (int *)(b + oCellOf), (int *)(b + oStart),
                                     (int *)(b + oItems));
    MCK(cudaGetLastError());
    const int blocks = (nq + 127) / 128;
    k_area_count<<<blocks, 128, 0, st>>>(dKeys, dStart, dItems, frame->min_x, frame->min_y, invW, invH, (const float *)(b + oQX),
                                         (const float *)(b + oQY), (const float *)(b + oQR), (const int *)(b + oQL0),
                                         (const int *)(b + oQL1), nq, (int *)(b + oCount));
    MCK(cudaGetLastError());
    k_area_scan<<<1, 1024, 0, st>>>((const int *)(b + oCount), nq, (int *)(b + oOff));
    MCK(cudaGetLastError());
    k_area_fill<<<blocks, 128, 0, st>>>(dKeys, (const uint4 *)(b + oDesc), dStart, dItems, frame->min_x, frame->min_y, invW, invH,
                                        q_desc ? (const uint4 *)(b + oQDesc) : nullptr, (const float *)(b + oQX),
                                        (const float *)(b + oQY), (const float *)(b + oQR), (const int *)(b + oQL0),
                                        (const int *)(b + oQL1), nq, (const int *)(b + oOff), cap, (int *)(b + oInd), (int *)(b + oDist));
    MCK(cudaGetLastError());
    MCK(cudaMemcpyAsync(hb + oOff, b + oOff, (size_t)(nq + 1) * 4, cudaMemcpyDeviceToHost, st));
    MCK(cudaStreamSynchronize(st));
    const int total = ((const int *)(hb + oOff))[nq];
    *n_entries = total;
    if (total > cap) {
        char msg[128];
        snprintf(msg, sizeof msg, "%d features found, the output arrays hold %d", total, cap);
        return mfail(m, ORBX_ERR_CAPACITY, msg);
    }
    memcpy(offsets, hb + oOff, (size_t)(nq + 1) * 4);
    if (total > 0) {
        MCK(cudaMemcpyAsync(hb + oInd, b + oInd, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
        if (q_desc) MCK(cudaMemcpyAsync(hb + oDist, b + oDist, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
        MCK(cudaStreamSynchronize(st));
        memcpy(indices, hb + oInd, (size_t)total * 4);
        if (q_desc) memcpy(dist, hb + oDist, (size_t)total * 4);
    }
    (void)oRes;
    return ORBX_OK;
}


int orbm_assign_grid(orbm_matcher *m, const orbx_keypoint *keys, int n, float min_x, float min_y, float max_x, float max_y,
                     int32_t *cell_start, int32_t *cell_items)
{
    if (!m) return ORBX_ERR_ARG;
    if (!cell_start || n < 0 || (n > 0 && (!keys || !cell_items)) || !(max_x > min_x) || !(max_y > min_y))
        return mfail(m, ORBX_ERR_ARG, "bad argument");
    for (int c = 0; c <= FG_CELLS; c++) cell_start[c] = 0;
    if (n == 0) return ORBX_OK;
    MCK(cudaSetDevice(m->device));
    const float invW = (float)FG_COLS / (max_x - min_x), invH = (float)FG_ROWS / (max_y - min_y);
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t o = 0;
    const size_t oKeys = o;    o += al((size_t)n * sizeof(orbx_keypoint));
    const size_t oCellOf = o;  o += al((size_t)n * 4);
    const size_t oStart = o;   o += al((size_t)(FG_CELLS + 1) * 4);
    const size_t oItems = o;   o += al((size_t)n * 4);
    if (o > m->spCap) {
        if (m->spBuf) cudaFree(m->spBuf);
        if (m->spHost) cudaFreeHost(m->spHost);
        m->spBuf = nullptr; m->spHost = nullptr; m->spCap = 0;
        MCK(cudaMalloc((void **)&m->spBuf, o));
        MCK(cudaMallocHost((void **)&m->spHost, o));
        m->spCap = o;
    }
    uint8_t *b = m->spBuf, *hb = m->spHost;
    cudaStream_t st = m->stream;
    memcpy(hb + oKeys, keys, (size_t)n * sizeof(orbx_keypoint));
    MCK(cudaMemcpyAsync(b + oKeys, hb + oKeys, (size_t)n * sizeof(orbx_keypoint), cudaMemcpyHostToDevice, st));
    k_frame_grid<<<1, 1024, 0, st>>>((const orbx_keypoint *)(b + oKeys), n, min_x, min_y, invW, invH, (int *)(b + oCellOf),
                                     (int *)(b + oStart), (int *)(b + oItems));
    MCK(cudaGetLastError());
    MCK(cudaMemcpyAsync(hb + oStart, b + oStart, (oItems - oStart) + (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    MCK(cudaStreamSynchronize(st));
    memcpy(cell_start, hb + oStart, (size_t)(FG_CELLS + 1) * 4);
    memcpy(cell_items, hb + oItems, (size_t)cell_start[FG_CELLS] * 4);
    return ORBX_OK;
}

// ---- several GPUs behind one matcher handle (one process): queries sharded, train replicated, k_knn2_sharded with peer stores
struct orbm_multi {
    std::vector<orbm_matcher *> m;
    int maxQ = 0, nt = -1;
    std::string err;
};

static int mmfail(orbm_multi *mm, int code, const std::string &msg) { if (mm) mm->err = msg; return code; }

int orbm_multi_create(const int *devices, int n_devices, int max_queries, int max_train, orbm_multi **out)
{
    if (!out || !devices || n_devices < 1 || n_devices > KNN_MAX_RANKS || max_queries < 1 || max_train < 1) return ORBX_ERR_ARG;
    orbm_multi *mm = new (std::nothrow) orbm_multi();
    if (!mm) return ORBX_ERR_NOMEM;
    *out = mm;
    mm->maxQ = max_queries;
    const int block = (max_queries + n_devices - 1) / n_devices;
    for (int g = 0; g < n_devices; g++) {
        orbm_matcher *m = nullptr;
        int rc = orbm_create(devices[g], block, max_train, &m);
        if (m) mm->m.push_back(m);
        if (rc != ORBX_OK) return mmfail(mm, rc, std::string("device ") + std::to_string(devices[g]) + ": " + orbm_last_error(m));
        rc = orbm_window_create(m, max_queries, n_devices, g, nullptr);
        if (rc != ORBX_OK) return mmfail(mm, rc, orbm_last_error(m));
    }
    for (int g = 0; g < n_devices; g++)
        for (int r = 0; r < n_devices; r++)
            if (r != g) {
                const int rc = orbm_window_attach_peer(mm->m[g], r, mm->m[r]);
                if (rc != ORBX_OK) return mmfail(mm, rc, orbm_last_error(mm->m[g]));
            }
    return ORBX_OK;
}

void orbm_multi_destroy(orbm_multi *mm)
{
    if (!mm) return;
    for (orbm_matcher *m : mm->m) if (m && m->stream) { cudaSetDevice(m->device); cudaStreamSynchronize(m->stream); }
    for (orbm_matcher *m : mm->m) orbm_destroy(m);
    delete mm;
}

const char *orbm_multi_last_error(const orbm_multi *mm) { return mm ? mm->err.c_str() : "null handle"; }
int orbm_multi_devices(const orbm_multi *mm) { return mm ? (int)mm->m.size() : ORBX_ERR_ARG; }

int orbm_multi_set_train(orbm_multi *mm, const uint8_t *t, int nt)
{
    if (!mm) return ORBX_ERR_ARG;
    for (orbm_matcher *m : mm->m) {
        const int rc = orbm_set_train(m, t, nt);
        if (rc != ORBX_OK) return mmfail(mm, rc, orbm_last_error(m));
    }
    mm->nt = nt;
    return ORBX_OK;
}

int orbm_multi_knn2(orbm_multi *mm, const uint8_t *q, int nq, int32_t *idx, int32_t *d1, int32_t *d2)
{
    if (!mm) return ORBX_ERR_ARG;
    if (mm->nt < 1) return mmfail(mm, ORBX_ERR_ARG, "no resident train set (orbm_multi_set_train)");
    if (!q || !idx || !d1 || !d2 || nq < 1 || nq > mm->maxQ) return mmfail(mm, ORBX_ERR_ARG, "bad query block");
    const int G = (int)mm->m.size();
    const int block = (nq + G - 1) / G;
    // every GPU gets its block of queries and launches; the kernels exchange the records among themselves
    for (int g = 0; g < G; g++) {
        orbm_matcher *m = mm->m[g];
        const int lo = std::min(g * block, nq), hi = std::min(lo + block, nq);
        cudaError_t e = cudaSetDevice(m->device);
        if (e == cudaSuccess && hi > lo) e = cudaMemcpyAsync(m->dQ, q + (size_t)lo * 32, (size_t)(hi - lo) * 32, cudaMemcpyHostToDevice, m->stream);
        if (e != cudaSuccess) return mmfail(mm, ORBX_ERR_CUDA, cudaGetErrorString(e));
        const int rc = orbm_knn2_sharded(m, m->dQ, hi - lo, lo, m->dT, mm->nt, nullptr);
        if (rc != ORBX_OK) return mmfail(mm, rc, orbm_last_error(m));
    }
    // the complete result is in every GPU's window; the host reads GPU 0's
    orbm_matcher *m0 = mm->m[0];
    const int32_t *rec = nullptr;
    int rc = orbm_window_records(m0, &rec);
    if (rc != ORBX_OK) return mmfail(mm, rc, orbm_last_error(m0));
    std::vector<int4> host((size_t)nq);
    cudaError_t e = cudaSetDevice(m0->device);
    if (e == cudaSuccess) e = cudaMemcpyAsync(host.data(), rec, (size_t)nq * sizeof(int4), cudaMemcpyDeviceToHost, m0->stream);
    if (e != cudaSuccess) return mmfail(mm, ORBX_ERR_CUDA, cudaGetErrorString(e));
    for (int g = 0; g < G; g++) {
        rc = orbm_window_status(mm->m[g], nullptr);
        if (rc != ORBX_OK) return mmfail(mm, rc, orbm_last_error(mm->m[g]));
    }
    for (int i = 0; i < nq; i++) { idx[i] = host[i].x; d1[i] = host[i].y; d2[i] = host[i].z; }
    return ORBX_OK;
}

/* the matcher of device slot g (its window holds the complete result after orbm_multi_knn2) */
orbm_matcher *orbm_multi_matcher(orbm_multi *mm, int g) { return mm && g >= 0 && g < (int)mm->m.size() ? mm->m[g] : nullptr; }

} // extern "C"
