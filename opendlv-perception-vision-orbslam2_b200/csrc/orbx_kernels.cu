// orbx_kernels.cu -- sm_100a kernels of the ORB extractor path.
//
// Each kernel names the reference lines (under /root/reference) whose results it reproduces
// bit-for-bit; the arithmetic models of the OpenCV primitives are SURVEY.md Appendix A.
//
//   k_copy_level0     level-0 copy of ComputePyramid            orbextractor.cpp:654-678
//   k_resize          cv::resize(INTER_LINEAR) level l <- l-1   orbextractor.cpp:666      (A.1)
//   k_fast_cells      gridded FAST-9 + NMS + threshold fallback orbextractor.cpp:906-970  (A.3)
//   k_octree          DistributeOctTree + DivideNode            orbextractor.cpp:680-904, :72-128 (A.5)
//   k_blur            cv::GaussianBlur 7x7 sigma 2 REFLECT_101  orbextractor.cpp:621-622  (A.4)
//   k_describe        IC_Angle + rBRIEF + keypoint assembly     orbextractor.cpp:136-211, :978-988, :631-639
#include "orbx_internal.h"
#include "orbx_kernels.h"

#include <cuda_runtime.h>

// rBRIEF sampling pattern, 512 (x,y) points (data; same table as orbextractor.cpp:215-473)
__device__ const int8_t d_pattern[1024] = {
#include "orb_pattern.inc"
};

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// In-place exclusive scan of vals[0..n) by the whole block; returns the total.
// scratch: >= 33 ints of shared memory.  Contains __syncthreads(); call uniformly.
__device__ int block_excl_scan(int *vals, int n, int *scratch)
{
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n + T - 1) / T;
    int beg = tid * per; if (beg > n) beg = n;
    int end = beg + per; if (end > n) end = n;
    int sum = 0;
    for (int i = beg; i < end; i++) sum += vals[i];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) scratch[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int nw = (T + 31) >> 5;
        int v = lane < nw ? scratch[lane] : 0;
        int inc2 = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc2, o);
            if (lane >= o) inc2 += t;
        }
        if (lane < nw) scratch[lane] = inc2 - v; // exclusive warp offsets
        if (lane == 31) scratch[32] = inc2;      // total
    }
    __syncthreads();
    int run = incl - sum + scratch[warp];
    const int total = scratch[32];
    for (int i = beg; i < end; i++) { int t = vals[i]; vals[i] = run; run += t; }
    __syncthreads();
    return total;
}

// ------------------------------------------------------------------------------------------
// level 0: copy the caller's frame into the pyramid slab (ComputePyramid level 0; the
// REFLECT_101 border of orbextractor.cpp:673 is never read by the extractor -- SURVEY A.2)
// ------------------------------------------------------------------------------------------
__global__ void k_copy_level0(const uint8_t *__restrict__ src, size_t frameStride, size_t srcPitch,
                              uint8_t *__restrict__ pyr, long long slab, int off, int pitch, int w, int h)
{
    const int f = blockIdx.z;
    const int y = blockIdx.y;
    const uint8_t *s = src + (size_t)f * frameStride + (size_t)y * srcPitch;
    uint8_t *d = pyr + (size_t)f * slab + off + (size_t)y * pitch;
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (x4 >= w) return;
    if (((uintptr_t)(s + x4) & 15) == 0 && x4 + 16 <= w) {
        *(uint4 *)(d + x4) = __ldg((const uint4 *)(s + x4));
    } else {
        for (int k = 0; k < 16 && x4 + k < w; k++) d[x4 + k] = s[x4 + k];
    }
}

void launch_copy_level0(const uint8_t *src, size_t frameStride, size_t srcPitch, uint8_t *pyr,
                        const OrbxLayout &L, int batch, cudaStream_t st)
{
    const OrbxLevel &l0 = L.lv[0];
    dim3 block(64);
    dim3 grid((l0.w + 16 * 64 - 1) / (16 * 64), l0.h, batch);
    k_copy_level0<<<grid, block, 0, st>>>(src, frameStride, srcPitch, pyr, L.slab, l0.off, l0.pitch, l0.w, l0.h);
}

// ------------------------------------------------------------------------------------------
// pyramid resize, 11-bit fixed-point bilinear (A.1).  Each thread produces 4 horizontally
// adjacent output pixels and stores them as one 32-bit word.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_resize(uint8_t *__restrict__ pyr, long long slab, int srcOff, int srcPitch, int sw, int sh,
         int dstOff, int dstPitch, int dw, int dh,
         const OrbxRTab *__restrict__ xtab, const OrbxRTab *__restrict__ ytab)
{
    const int f = blockIdx.z;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (y >= dh || x0 >= dw) return;
    const uint8_t *src = pyr + (size_t)f * slab + srcOff;
    uint8_t *dst = pyr + (size_t)f * slab + dstOff;
    const OrbxRTab ty = ytab[y];
    const int sy0 = ty.ofs, sy1 = min(sy0 + 1, sh - 1);
    const uint8_t *r0 = src + (size_t)sy0 * srcPitch, *r1 = src + (size_t)sy1 * srcPitch;
    const int b0 = ty.c0, b1 = ty.c1;
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int x = x0 + k;
        if (x < dw) {
            const OrbxRTab tx = xtab[x];
            const int sx0 = tx.ofs, sx1 = min(sx0 + 1, sw - 1);
            const int h0 = r0[sx0] * tx.c0 + r0[sx1] * tx.c1;
            const int h1 = r1[sx0] * tx.c0 + r1[sx1] * tx.c1;
            const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
            out |= (uint32_t)(v & 255) << (8 * k);
        }
    }
    *(uint32_t *)(dst + (size_t)y * dstPitch + x0) = out; // pitch is a multiple of 128: in-row padding absorbs the tail
}

void launch_resize(uint8_t *pyr, const OrbxLayout &L, int level, const OrbxRTab *tabs, int batch, cudaStream_t st)
{
    const OrbxLevel &s = L.lv[level - 1], &d = L.lv[level];
    dim3 block(32, 8);
    dim3 grid((d.w + 127) / 128, (d.h + 7) / 8, batch);
    k_resize<<<grid, block, 0, st>>>(pyr, L.slab, s.off, s.pitch, s.w, s.h, d.off, d.pitch, d.w, d.h,
                                     tabs + d.xtabOff, tabs + d.ytabOff);
}

// ------------------------------------------------------------------------------------------
// 7x7 Gaussian blur, separable integer taps, REFLECT_101 (A.4).
// CTA tile: 64x32 outputs; shared-memory stage with a 3-pixel halo.
// ------------------------------------------------------------------------------------------
#define BL_TW 64
#define BL_TH 32
struct BlurTaps { int t[7]; };

__global__ void __launch_bounds__(256)
k_blur(const uint8_t *__restrict__ pyr, uint8_t *__restrict__ blur, long long slab, int off, int pitch,
       int w, int h, BlurTaps taps)
{
    __shared__ uint8_t raw[(BL_TH + 6) * (BL_TW + 8)];
    __shared__ uint16_t hb[(BL_TH + 6) * BL_TW];
    const int f = blockIdx.z;
    const int tx0 = blockIdx.x * BL_TW, ty0 = blockIdx.y * BL_TH;
    const uint8_t *src = pyr + (size_t)f * slab + off;
    uint8_t *dst = blur + (size_t)f * slab + off;
    const int tid = threadIdx.x;
    const int RW = BL_TW + 6, RP = BL_TW + 8;
    for (int i = tid; i < (BL_TH + 6) * RW; i += 256) {
        const int r = i / RW, c = i - r * RW;
        const int gy = reflect101(ty0 + r - 3, h), gx = reflect101(tx0 + c - 3, w);
        raw[r * RP + c] = src[(size_t)gy * pitch + gx];
    }
    __syncthreads();
    for (int i = tid; i < (BL_TH + 6) * BL_TW; i += 256) {
        const int r = i / BL_TW, c = i - r * BL_TW;
        const uint8_t *p = &raw[r * RP + c];
        uint32_t acc = 0;
#pragma unroll
        for (int k = 0; k < 7; k++) acc += (uint32_t)taps.t[k] * p[k];
        hb[r * BL_TW + c] = (uint16_t)acc;
    }
    __syncthreads();
    for (int i = tid; i < BL_TH * BL_TW; i += 256) {
        const int r = i / BL_TW, c = i - r * BL_TW;
        const int gy = ty0 + r, gx = tx0 + c;
        if (gy < h && gx < w) {
            uint32_t acc = 0;
#pragma unroll
            for (int k = 0; k < 7; k++) acc += (uint32_t)taps.t[k] * hb[(r + k) * BL_TW + c];
            uint32_t v = (acc + 32768u) >> 16;
            dst[(size_t)gy * pitch + gx] = (uint8_t)(v > 255u ? 255u : v);
        }
    }
}

void launch_blur(const uint8_t *pyr, uint8_t *blur, const OrbxLayout &L, int level, const int taps[7], int batch, cudaStream_t st)
{
    const OrbxLevel &l = L.lv[level];
    BlurTaps t;
    for (int k = 0; k < 7; k++) t.t[k] = taps[k];
    dim3 grid((l.w + BL_TW - 1) / BL_TW, (l.h + BL_TH - 1) / BL_TH, batch);
    k_blur<<<grid, 256, 0, st>>>(pyr, blur, L.slab, l.off, l.pitch, l.w, l.h, t);
}

// ------------------------------------------------------------------------------------------
// Gridded FAST-9/16 with 3x3 NMS per cell and the ini/min threshold fallback (A.3).
// One CTA per (cell, frame).  Output is not a keypoint list: DistributeOctTree in this fork only
// ever splits along y inside fixed x-strips (A.5), so all it needs per (strip, row) is the number
// of candidates and the best candidate (max response, first in the reference's emission order).
// Both are accumulated here with atomics:
//   cnt [frame][rowBase + strip*H + y]  += 1
//   best[frame][rowBase + strip*H + y]   = max(score<<56 | ~order<<28 | x<<14 | y)
// `order` = (cell index << 12 | yIn << 6 | xIn) is the position in the reference's emission
// order (cells row-major, raster inside a cell, orbextractor.cpp:930-968).
// ------------------------------------------------------------------------------------------
#define FW_P 72   // shared window pitch (bytes)
#define FS_P 64   // shared score-map pitch

__device__ __forceinline__ int fast_score16(const int (&d)[16])
{
    int mn1[16], mx1[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { mn1[k] = min(d[k], d[(k + 1) & 15]); mx1[k] = max(d[k], d[(k + 1) & 15]); }
    int mn2[16], mx2[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { mn2[k] = min(mn1[k], mn1[(k + 2) & 15]); mx2[k] = max(mx1[k], mx1[(k + 2) & 15]); }
    int a = -256, b = 256;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        // window of 9 starting at k: [k..k+3] U [k+4..k+7] U {k+8}
        int mn = min(min(mn2[k], mn2[(k + 4) & 15]), d[(k + 8) & 15]);
        int mx = max(max(mx2[k], mx2[(k + 4) & 15]), d[(k + 8) & 15]);
        a = max(a, mn);
        b = min(b, mx);
    }
    return max(a, -b) - 1;
}

__global__ void __launch_bounds__(128)
k_fast_cells(const uint8_t *__restrict__ pyr, const __grid_constant__ OrbxLayout L,
             const OrbxCell *__restrict__ cells, uint32_t *__restrict__ cnt,
             unsigned long long *__restrict__ best, OrbxDbgCand *__restrict__ dbg,
             int *__restrict__ dbgCount, int dbgCap)
{
    __shared__ __align__(16) uint8_t win[66 * FW_P];
    __shared__ __align__(16) uint8_t smap[62 * FS_P];
    __shared__ uint16_t cand[60 * 60];
    __shared__ int ncand;

    const OrbxCell cell = cells[blockIdx.x];
    const int frame = blockIdx.y;
    const OrbxLevel &lv = L.lv[cell.level];
    const int wEff = (int)cell.w - 6, hEff = (int)cell.h - 6;
    if (wEff <= 0 || hEff <= 0) return;
    const int tid = threadIdx.x, lane = tid & 31;
    const int th = L.minTh;

    // ---- stage the window: aligned 32-bit loads of each row span
    const uint8_t *base = pyr + (size_t)frame * L.slab + lv.off;
    const int xa = cell.x0 & ~3, shift = cell.x0 - xa;
    const int nw = (shift + cell.w + 3) >> 2;
    for (int i = tid; i < cell.h * nw; i += 128) {
        const int r = i / nw, k = i - r * nw;
        const uint32_t v = __ldg((const uint32_t *)(base + (size_t)(cell.y0 + r) * lv.pitch + xa) + k);
        ((uint32_t *)win)[r * (FW_P / 4) + k] = v;
    }
    for (int i = tid; i < 62 * FS_P / 4; i += 128) ((uint32_t *)smap)[i] = 0;
    if (tid == 0) ncand = 0;
    __syncthreads();

    // ---- stage 1: compass quick-reject at the lower threshold, warp-ballot compaction
    const int total = wEff * hEff;
    for (int p0 = 0; p0 < total; p0 += 128) {
        const int p = p0 + tid;
        bool pass = false;
        int yIn = 0, xIn = 0;
        if (p < total) {
            yIn = p / wEff; xIn = p - yIn * wEff;
            const uint8_t *c = &win[(yIn + 3) * FW_P + shift + xIn + 3];
            const int v = c[0], hi = v + th, lo = v - th;
            const int a0 = c[3 * FW_P], a8 = c[-3 * FW_P], a4 = c[3], a12 = c[-3];
            const bool br = ((a0 > hi) | (a8 > hi)) & ((a4 > hi) | (a12 > hi));
            const bool dk = ((a0 < lo) | (a8 < lo)) & ((a4 < lo) | (a12 < lo));
            pass = br | dk;
        }
        const unsigned m = __ballot_sync(0xffffffffu, pass);
        if (m) {
            int b = 0;
            if (lane == 0) b = atomicAdd(&ncand, __popc(m));
            b = __shfl_sync(0xffffffffu, b, 0);
            if (pass) cand[b + __popc(m & ((1u << lane) - 1u))] = (uint16_t)(yIn << 6 | xIn);
        }
    }
    __syncthreads();
    const int nc = ncand;

    // ---- stage 2: full 16-pixel ring test and corner score for the survivors
    for (int i = tid; i < nc; i += 128) {
        const int yIn = cand[i] >> 6, xIn = cand[i] & 63;
        const uint8_t *c = &win[(yIn + 3) * FW_P + shift + xIn + 3];
        const int v = c[0];
        int d[16];
        d[0] = v - c[3 * FW_P];       d[1] = v - c[3 * FW_P + 1];   d[2] = v - c[2 * FW_P + 2];   d[3] = v - c[FW_P + 3];
        d[4] = v - c[3];              d[5] = v - c[-FW_P + 3];      d[6] = v - c[-2 * FW_P + 2];  d[7] = v - c[-3 * FW_P + 1];
        d[8] = v - c[-3 * FW_P];      d[9] = v - c[-3 * FW_P - 1];  d[10] = v - c[-2 * FW_P - 2]; d[11] = v - c[-FW_P - 3];
        d[12] = v - c[-3];            d[13] = v - c[FW_P - 3];      d[14] = v - c[2 * FW_P - 2];  d[15] = v - c[3 * FW_P - 1];
        unsigned mb = 0, md = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) { mb |= (unsigned)(d[k] > th) << k; md |= (unsigned)(d[k] < -th) << k; }
        mb |= mb << 16; md |= md << 16;
        unsigned r = mb & (mb >> 1); r &= r >> 2; r &= r >> 4; r &= mb >> 8;
        unsigned q = md & (md >> 1); q &= q >> 2; q &= q >> 4; q &= md >> 8;
        if (r | q) smap[(yIn + 1) * FS_P + xIn + 1] = (uint8_t)fast_score16(d);
    }
    __syncthreads();

    // ---- stage 3: 3x3 non-maximum suppression (strict >, outside the cell interior counts as 0)
    int any = 0;
    for (int i = tid; i < nc; i += 128) {
        const int yIn = cand[i] >> 6, xIn = cand[i] & 63;
        const uint8_t *s = &smap[(yIn + 1) * FS_P + xIn + 1];
        const int v = s[0];
        if (v && v > s[-1] && v > s[1] && v > s[-FS_P - 1] && v > s[-FS_P] && v > s[-FS_P + 1] &&
            v > s[FS_P - 1] && v > s[FS_P] && v > s[FS_P + 1]) {
            cand[i] |= 0x8000;
            any |= (v >= L.iniTh);
        }
    }
    // per-cell threshold fallback, orbextractor.cpp:950-957: the ini-threshold result is used iff
    // it is non-empty after NMS; NMS(ini) == {k in NMS(min) : score >= ini}
    const int haveIni = __syncthreads_or(any);

    // ---- stage 4: emit into the per-(strip,row) summaries
    const int cellIdx = cell.orderBase >> 12;
    const int ci = cellIdx / lv.nCols, cj = cellIdx - ci * lv.nCols;
    uint32_t *cntF = cnt + (size_t)frame * L.rowsPerFrame + lv.rowBase;
    unsigned long long *bestF = best + (size_t)frame * L.rowsPerFrame + lv.rowBase;
    for (int i = tid; i < nc; i += 128) {
        const int e = cand[i];
        if (!(e & 0x8000)) continue;
        const int yIn = (e >> 6) & 63, xIn = e & 63;
        const int s = smap[(yIn + 1) * FS_P + xIn + 1];
        if (haveIni && s < L.iniTh) continue;
        const int xr = cj * lv.wCell + 3 + xIn, yr = ci * lv.hCell + 3 + yIn; // relative to (16,16), :963-964
        const int strip = xr / lv.hX;                                         // :710
        const int row = strip * lv.H + yr;
        const unsigned order = cell.orderBase | (unsigned)(yIn << 6 | xIn);
        const unsigned long long key = ((unsigned long long)s << 56) |
                                       ((unsigned long long)(0x0fffffffu - order) << 28) |
                                       ((unsigned long long)xr << 14) | (unsigned long long)yr;
        atomicAdd(&cntF[row], 1u);
        atomicMax(&bestF[row], key);
        if (dbg) {
            const int slot = frame * L.nlevels + cell.level;
            const int pos = atomicAdd(&dbgCount[slot], 1);
            if (pos < dbgCap) {
                OrbxDbgCand c; c.xy = xr | (yr << 16); c.score = s;
                dbg[(size_t)slot * dbgCap + pos] = c;
            }
        }
    }
}

void launch_fast(const uint8_t *pyr, const OrbxLayout &L, const OrbxCell *cells, uint32_t *cnt,
                 unsigned long long *best, OrbxDbgCand *dbg, int *dbgCount, int dbgCap, int batch, cudaStream_t st)
{
    dim3 grid(L.totalCells, batch);
    k_fast_cells<<<grid, 128, 0, st>>>(pyr, L, cells, cnt, best, dbg, dbgCount, dbgCap);
}

// ------------------------------------------------------------------------------------------
// DistributeOctTree (A.5): one CTA per (level, frame) replays the reference's list algorithm
// level-synchronously.  A node is (strip, [y0,y1)); its size is a difference of the per-strip
// prefix sums P of the row counts, so dividing a node is O(1).
//   main pass   : every node with >1 point is divided; children go to the list front in the
//                 order n2, n4 (=> n4 first), parents processed front to back     (:735-808)
//   priority    : nodes sorted by (size, creation) descending, divided until size >= N (:814-878)
//   output      : per node the max-response point, first in emission order on ties  (:885-901)
// Under a monotone allocator "higher address" == "created later"; all expandable nodes of a
// round were created in the previous round and sit at the list front in reverse creation order,
// so "newest first" == "lowest list position first".
// ------------------------------------------------------------------------------------------
#define OCT_T 256
__device__ __forceinline__ uint32_t node_pack(int s, int y0, int y1) { return (uint32_t)s << 26 | (uint32_t)y0 << 13 | (uint32_t)y1; }

struct OctSh {
    int n, rec, jstar, totC, R;
    int scan[34];
};

__global__ void __launch_bounds__(OCT_T)
k_octree(const __grid_constant__ OrbxLayout L, uint32_t *__restrict__ cnt,
         const unsigned long long *__restrict__ best, int2 *__restrict__ slots,
         int *__restrict__ lvlCount, int maxRows, int maxNodes, int pow2Nodes)
{
    extern __shared__ __align__(16) unsigned char sm_raw[];
    __shared__ OctSh sh;
    const int level = blockIdx.x, frame = blockIdx.y;
    const OrbxLevel &lv = L.lv[level];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = lv.H, nR = lv.nIni * H, N = lv.quota;

    unsigned long long *keys = (unsigned long long *)sm_raw;         // pow2Nodes
    int *P = (int *)(keys + pow2Nodes);                              // maxRows + 1
    uint32_t *cur = (uint32_t *)(P + maxRows + 1);                   // maxNodes
    uint32_t *nxt = cur + maxNodes;
    int *a = (int *)(nxt + maxNodes), *b = a + maxNodes, *c = b + maxNodes, *d = c + maxNodes;

    const uint32_t *cntF = cnt + (size_t)frame * L.rowsPerFrame + lv.rowBase;
    const unsigned long long *bestF = best + (size_t)frame * L.rowsPerFrame + lv.rowBase;

    for (int i = tid; i < nR; i += OCT_T) P[i] = (int)cntF[i];
    __syncthreads();
    const int totalCand = block_excl_scan(P, nR, sh.scan);
    if (tid == 0) {
        P[nR] = totalCand;
        int n = 0;
        for (int s = 0; s < lv.nIni; s++)           // :691-726 initial strips, empty ones erased
            if (P[(s + 1) * H] - P[s * H] > 0) cur[n++] = node_pack(s, 0, H);
        sh.n = n;
    }
    __syncthreads();
    int n = sh.n;

#define NODE_DECODE(node) const int s_ = (node) >> 26, y0_ = ((node) >> 13) & 8191, y1_ = (node)&8191; const int pb_ = s_ * H
#define NODE_SIZE() (P[pb_ + y1_] - P[pb_ + y0_])

    bool finish = (n == 0);
    while (!finish) {
        const int prevSize = n;
        // ---------------- main pass
        int myrec = 0;
        if (tid == 0) sh.rec = 0;
        for (int i = tid; i < n; i += OCT_T) {
            const uint32_t node = cur[i];
            NODE_DECODE(node);
            const int sz = NODE_SIZE();
            if (sz > 1) {
                const int mid = y0_ + ((y1_ - y0_) >> 1);      // :75 integer halfY
                const int c2 = P[pb_ + mid] - P[pb_ + y0_], c4 = sz - c2;
                a[i] = (c2 > 0) + (c4 > 0); b[i] = 0;
                myrec += (c2 > 1) + (c4 > 1);
            } else { a[i] = 0; b[i] = 1; }
        }
        __syncthreads();
        if (myrec) atomicAdd(&sh.rec, myrec);
        const int totC = block_excl_scan(a, n, sh.scan);
        const int totN = block_excl_scan(b, n, sh.scan);
        for (int i = tid; i < n; i += OCT_T) {
            const uint32_t node = cur[i];
            NODE_DECODE(node);
            const int sz = NODE_SIZE();
            if (sz > 1) {
                const int mid = y0_ + ((y1_ - y0_) >> 1);
                const int c2 = P[pb_ + mid] - P[pb_ + y0_], c4 = sz - c2;
                int pos = totC - (a[i] + (c2 > 0) + (c4 > 0));
                if (c4 > 0) nxt[pos++] = node_pack(s_, mid, y1_);
                if (c2 > 0) nxt[pos] = node_pack(s_, y0_, mid);
            } else {
                nxt[totC + b[i]] = node;
            }
        }
        __syncthreads();
        { uint32_t *t = cur; cur = nxt; nxt = t; }
        n = totC + totN;
        const int nToExpand = sh.rec;
        if (n > maxNodes) n = maxNodes; // cannot happen for supported shapes (see orbx_api geometry checks)
        __syncthreads();

        if (n >= N || n == prevSize) {
            finish = true;
        } else if (n + nToExpand * 3 > N) {
            // ---------------- priority rounds
            while (!finish) {
                const int prev2 = n;
                int p2 = 2; while (p2 < n) p2 <<= 1;
                if (tid == 0) { sh.R = 0; }
                __syncthreads();
                int myR = 0;
                for (int i = tid; i < p2; i += OCT_T) {
                    unsigned long long key = 0;
                    if (i < n) {
                        const uint32_t node = cur[i];
                        NODE_DECODE(node);
                        const int sz = NODE_SIZE();
                        if (sz > 1) {
                            key = (unsigned long long)sz << 16 | (unsigned)(L.tieRule ? i : 0xffff - i);
                            myR++;
                        }
                    }
                    keys[i] = key;
                }
                if (myR) atomicAdd(&sh.R, myR);
                __syncthreads();
                const int R = sh.R;
                if (R == 0) { finish = true; break; }   // nothing expandable: size unchanged (:875)
                for (int k = 2; k <= p2; k <<= 1)
                    for (int j = k >> 1; j > 0; j >>= 1) {
                        for (int t = tid; t < p2; t += OCT_T) {
                            const int x = t ^ j;
                            if (x > t) {
                                const unsigned long long ka = keys[t], kb = keys[x];
                                const bool desc = (t & k) == 0;
                                if ((ka < kb) == desc) { keys[t] = kb; keys[x] = ka; }
                            }
                        }
                        __syncthreads();
                    }
                // sorted position j -> list index, children, gain
                for (int j = tid; j < R; j += OCT_T) {
                    const int low = (int)(keys[j] & 0xffff);
                    const int i = L.tieRule ? low : 0xffff - low;
                    const uint32_t node = cur[i];
                    NODE_DECODE(node);
                    const int sz = NODE_SIZE();
                    const int mid = y0_ + ((y1_ - y0_) >> 1);
                    const int c2 = P[pb_ + mid] - P[pb_ + y0_], c4 = sz - c2;
                    const int nch = (c2 > 0) + (c4 > 0);
                    a[j] = nch - 1; b[j] = nch;
                }
                if (tid == 0) sh.jstar = R - 1;
                __syncthreads();
                // keep the per-j child counts: scans run in place, so stash nch in d[]
                for (int j = tid; j < R; j += OCT_T) d[j] = b[j];
                __syncthreads();
                block_excl_scan(a, R, sh.scan);
                block_excl_scan(b, R, sh.scan);
                for (int j = tid; j < R; j += OCT_T)
                    if (n + a[j] + (d[j] - 1) >= N) atomicMin(&sh.jstar, j);  // :871 break once size >= N
                __syncthreads();
                const int jstar = sh.jstar;
                if (tid == 0) sh.totC = b[jstar] + d[jstar];
                for (int i = tid; i < n; i += OCT_T) c[i] = 1;
                __syncthreads();
                const int totC2 = sh.totC;
                for (int j = tid; j <= jstar; j += OCT_T) {
                    const int low = (int)(keys[j] & 0xffff);
                    const int i = L.tieRule ? low : 0xffff - low;
                    c[i] = 0;
                    const uint32_t node = cur[i];
                    NODE_DECODE(node);
                    const int sz = NODE_SIZE();
                    const int mid = y0_ + ((y1_ - y0_) >> 1);
                    const int c2 = P[pb_ + mid] - P[pb_ + y0_], c4 = sz - c2;
                    int pos = totC2 - (b[j] + d[j]);      // later-processed parents' children sit nearer the front
                    if (c4 > 0) nxt[pos++] = node_pack(s_, mid, y1_);
                    if (c2 > 0) nxt[pos] = node_pack(s_, y0_, mid);
                }
                __syncthreads();
                for (int i = tid; i < n; i += OCT_T) a[i] = c[i];
                __syncthreads();
                const int totU = block_excl_scan(c, n, sh.scan);
                for (int i = tid; i < n; i += OCT_T)
                    if (a[i]) nxt[totC2 + c[i]] = cur[i];
                __syncthreads();
                { uint32_t *t = cur; cur = nxt; nxt = t; }
                n = totC2 + totU;
                if (n > maxNodes) n = maxNodes;
                if (n >= N || n == prev2) finish = true;
                __syncthreads();
            }
        }
    }

    // ---------------- output: best point of each node, list order (:885-901)
    int2 *out = slots + (size_t)frame * L.slotsPerFrame + lv.slotBase;
    const int nOut = n < lv.slotCap ? n : lv.slotCap;
    for (int i = warp; i < nOut; i += OCT_T / 32) {
        const uint32_t node = cur[i];
        NODE_DECODE(node);
        unsigned long long k = 0;
        for (int y = y0_ + lane; y < y1_; y += 32) {
            const unsigned long long v = bestF[pb_ + y];
            k = v > k ? v : k;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long v = __shfl_xor_sync(0xffffffffu, k, o);
            k = v > k ? v : k;
        }
        if (lane == 0) {
            const int x = (int)((k >> 14) & 0x3fff), y = (int)(k & 0x3fff), sc = (int)(k >> 56);
            out[i] = make_int2(x | (y << 16), sc);
        }
    }
    if (tid == 0) lvlCount[frame * L.nlevels + level] = nOut;
#undef NODE_DECODE
#undef NODE_SIZE
}

size_t octree_smem_bytes(int maxRows, int maxNodes, int pow2Nodes)
{
    return (size_t)pow2Nodes * 8 + (size_t)(maxRows + 1) * 4 + (size_t)maxNodes * 4 * 6 + 16;
}

cudaError_t launch_octree(const OrbxLayout &L, uint32_t *cnt, const unsigned long long *best, int2 *slots,
                          int *lvlCount, int maxRows, int maxNodes, int pow2Nodes, int batch, cudaStream_t st)
{
    const size_t smem = octree_smem_bytes(maxRows, maxNodes, pow2Nodes);
    if (smem > 48 * 1024) { // opt in per call: the attribute is per device and this is a cheap host-side set
        cudaError_t e = cudaFuncSetAttribute(k_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid(L.nlevels, batch);
    k_octree<<<grid, OCT_T, smem, st>>>(L, cnt, best, slots, lvlCount, maxRows, maxNodes, pow2Nodes);
    return cudaSuccess;
}

// ------------------------------------------------------------------------------------------
// Orientation + descriptor + record assembly: one warp per keypoint slot.
//   IC_Angle      orbextractor.cpp:136-163 on the UNBLURRED level, cv::fastAtan2 model A.6
//   rBRIEF        orbextractor.cpp:165-203 on the blurred level, float32 without FMA, half-even
//   assembly      orbextractor.cpp:978-988 and :631-639
// cos/sin of the reference are glibc's float overloads (orbextractor.cpp:169); glibc computes
// them in double from a fixed polynomial (sincosf table) -- restated here so the rotated sample
// coordinates round identically.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    const float ax = fabsf(x), ay = fabsf(y);
    const float eps = (float)2.2204460492503131e-16;
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// glibc 2.39 sinf/cosf (sysdeps/ieee754/flt-32/s_sincosf.h, reduce_fast + sinf_poly) for 0 <= y < 120
__device__ __forceinline__ void glibc_sincosf(float y, float *sn, float *cs)
{
    const double hpi_inv = 0x1.45f306dc9c883p+23, hpi = 0x1.921fb54442d18p+0;
    const double C0 = 1.0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5, C3 = -0x1.6c087e89a359dp-10, C4 = 0x1.99343027bf8c3p-16;
    const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
    double x = (double)y;
    const double r = __dmul_rn(x, hpi_inv);
    const int n = ((int)r + 0x800000) >> 24;
    x = fma(-(double)n, hpi, x);
    const double sg = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    const double pc = (n & 2) ? -1.0 : 1.0;
    const double xs = __dmul_rn(x, sg), x2 = __dmul_rn(x, x);
    // sine polynomial of xs
    const double x3 = __dmul_rn(xs, x2);
    const double s1 = fma(x2, S3, S2);
    const double x7 = __dmul_rn(x3, x2);
    const double s = fma(x3, S1, xs);
    const float polyS = (float)fma(x7, s1, s);
    // cosine polynomial (coefficients negated in quadrants 2,3)
    const double x4 = __dmul_rn(x2, x2);
    const double c2 = fma(x2, pc * C4, pc * C3);
    const double c1 = fma(x2, pc * C1, pc * C0);
    const double x6 = __dmul_rn(x4, x2);
    const double c = fma(x4, pc * C2, c1);
    const float polyC = (float)fma(x6, c2, c);
    *sn = (n & 1) ? polyC : polyS;  // sinf: sinf_poly(x*s, x2, p, n)
    *cs = (n & 1) ? polyS : polyC;  // cosf: sinf_poly(x*s, x2, p, n ^ 1)
}

struct DescUmax { int u[16]; };

__global__ void __launch_bounds__(128)
k_describe(const uint8_t *__restrict__ pyr, const uint8_t *__restrict__ blur, const __grid_constant__ OrbxLayout L,
           const int2 *__restrict__ slots, const int *__restrict__ lvlCount, DescUmax um,
           orbx_keypoint_pod *__restrict__ kps, uint8_t *__restrict__ desc, int *__restrict__ counts)
{
    __shared__ int8_t pat[1024];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 256; i += 128) ((uint32_t *)pat)[i] = ((const uint32_t *)d_pattern)[i];
    __syncthreads();
    const int frame = blockIdx.y;
    const int slot = blockIdx.x * 4 + warp;
    if (slot >= L.slotsPerFrame) return;
    // level of this slot + exclusive prefix of the per-level counts
    const int *lc = lvlCount + frame * L.nlevels;
    int level = 0;
    for (int l = 1; l < L.nlevels; l++) if (slot >= L.lv[l].slotBase) level = l;
    const OrbxLevel &lv = L.lv[level];
    const int i = slot - lv.slotBase;
    int before = 0, total = 0;
    for (int l = 0; l < L.nlevels; l++) { const int cnt = lc[l]; if (l < level) before += cnt; total += cnt; }
    if (slot == 0 && lane == 0) counts[frame] = total;
    if (i >= lc[level]) return;
    const int2 sl = slots[(size_t)frame * L.slotsPerFrame + slot];
    const int cx = (sl.x & 0xffff) + ORBX_MINB, cy = (sl.x >> 16) + ORBX_MINB; // :984-985
    const size_t lbase = (size_t)frame * L.slab + lv.off;

    // ---- IC_Angle: lane = column u, loop over rows v (coalesced 31-byte row reads)
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int u = lane - 15, au = u < 0 ? -u : u;
        const uint8_t *cp = pyr + lbase + (size_t)cy * lv.pitch + cx + u;
        int colsum = 0;
#pragma unroll
        for (int v = -15; v <= 15; v++) {
            const int av = v < 0 ? -v : v;
            if (au <= um.u[av]) {
                const int val = cp[v * lv.pitch];
                colsum += val;
                m01 += v * val;
            }
        }
        m10 = u * colsum;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // ---- rBRIEF: lane = descriptor byte, 16 rotated samples each
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.0);
    float sa, ca;
    glibc_sincosf(__fmul_rn(angle, factorPI), &sa, &ca);
    const float a = ca, b = sa;
    const uint8_t *center = blur + lbase + (size_t)cy * lv.pitch + cx;
    const int8_t *pp = pat + lane * 32;
    int val = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int t[2];
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const float px = (float)pp[4 * k + 2 * e], py = (float)pp[4 * k + 2 * e + 1];
            const int row = __float2int_rn(__fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a)));
            const int col = __float2int_rn(__fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b)));
            t[e] = center[row * lv.pitch + col];
        }
        val |= (t[0] < t[1]) << k;
    }
    const int o = before + i;
    desc[((size_t)frame * L.kpStride + o) * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
        orbx_keypoint_pod kp;
        float fx = (float)cx, fy = (float)cy;
        if (level != 0) { fx = __fmul_rn(fx, lv.sf); fy = __fmul_rn(fy, lv.sf); }
        kp.x = fx; kp.y = fy; kp.size = (float)lv.kpSize; kp.angle = angle; kp.response = (float)sl.y;
        kp.octave = level; kp.class_id = -1;
        kps[(size_t)frame * L.kpStride + o] = kp;
    }
}

void launch_describe(const uint8_t *pyr, const uint8_t *blur, const OrbxLayout &L, const int2 *slots,
                     const int *lvlCount, const int umax[16], orbx_keypoint_pod *kps, uint8_t *desc, int *counts,
                     int batch, cudaStream_t st)
{
    DescUmax um;
    for (int k = 0; k < 16; k++) um.u[k] = umax[k];
    dim3 grid((L.slotsPerFrame + 3) / 4, batch);
    k_describe<<<grid, 128, 0, st>>>(pyr, blur, L, slots, lvlCount, um, kps, desc, counts);
}
