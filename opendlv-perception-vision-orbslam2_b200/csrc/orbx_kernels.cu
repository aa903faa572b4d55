// orbx_kernels.cu -- sm_100a kernels of the ORB extractor path.
//
// Each kernel names the reference lines (under /root/reference) whose results it reproduces
// bit-for-bit; the arithmetic models of the OpenCV primitives are SURVEY.md Appendix A.
//
//   k_copy_level0     level-0 copy of ComputePyramid            orbextractor.cpp:654-678
//   k_resize          cv::resize(INTER_LINEAR) level l <- l-1   orbextractor.cpp:666      (A.1)
//   k_fast_segs       gridded FAST-9 + NMS + threshold fallback orbextractor.cpp:906-970  (A.3)
//   k_octree          DistributeOctTree + DivideNode            orbextractor.cpp:680-904, :72-128 (A.5)
//   k_blur            cv::GaussianBlur 7x7 sigma 2 REFLECT_101  orbextractor.cpp:621-622  (A.4)
//   k_describe        IC_Angle + rBRIEF + keypoint assembly     orbextractor.cpp:136-211, :978-988, :631-639
#include "orbx_internal.h"
#include "orbx_kernels.h"
#include "orbx_smem_optin.h"

#include <cuda.h>
#include <algorithm>
#include <cstdlib>
#include <cuda_runtime.h>

// rBRIEF sampling pattern, 512 (x,y) points (data; same table as orbextractor.cpp:215-473)
__device__ __align__(16) const int8_t d_pattern[1024] = {
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
// REFLECT_101 border of orbextractor.cpp:673 is never read by the extractor -- SURVEY A.2).
// Source rows may start at any byte alignment (tightly packed 1241-wide frames): every thread
// assembles one aligned 16-byte destination chunk from aligned 32-bit source words with funnel
// shifts, so both sides move whole words.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_copy_level0(const uint8_t *__restrict__ src, size_t frameStride, size_t srcPitch, const uint8_t *srcEnd,
              uint8_t *__restrict__ pyr, long long slab, int off, int pitch, int w, int h)
{
    const int f = blockIdx.z;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (y >= h || x >= w) return;
    const uint8_t *s = src + (size_t)f * frameStride + (size_t)y * srcPitch + x;
    uint8_t *d = pyr + (size_t)f * slab + off + (size_t)y * pitch + x;
    const uintptr_t a = (uintptr_t)s;
    const uint32_t *s4 = (const uint32_t *)(a & ~(uintptr_t)3);
    if ((const uint8_t *)(s4 + 5) <= srcEnd) {
        const int sh = (int)(a & 3) * 8;
        const uint32_t w0 = __ldg(s4), w1 = __ldg(s4 + 1), w2 = __ldg(s4 + 2), w3 = __ldg(s4 + 3), w4 = __ldg(s4 + 4);
        uint4 o;
        o.x = __funnelshift_r(w0, w1, sh); o.y = __funnelshift_r(w1, w2, sh);
        o.z = __funnelshift_r(w2, w3, sh); o.w = __funnelshift_r(w3, w4, sh);
        *(uint4 *)d = o;    // bytes past w land in the row's pitch padding
    } else {
        for (int k = 0; k < 16 && x + k < w; k++) d[k] = s[k];
    }
}

void launch_copy_level0(const uint8_t *src, size_t frameStride, size_t srcPitch, uint8_t *pyr,
                        const OrbxLayout &L, int batch, cudaStream_t st)
{
    const OrbxLevel &l0 = L.lv[0];
    dim3 block(32, 4);
    dim3 grid((l0.w + 16 * 32 - 1) / (16 * 32), (l0.h + 3) / 4, batch);
    const uint8_t *srcEnd = src + (size_t)(batch - 1) * frameStride + (size_t)(l0.h - 1) * srcPitch + l0.w;
    k_copy_level0<<<grid, block, 0, st>>>(src, frameStride, srcPitch, srcEnd, pyr, L.slab, l0.off, l0.pitch, l0.w, l0.h);
}

// ---- TMA / mbarrier helpers (cp.async.bulk.tensor box loads completing on an mbarrier) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_load_tile_3d(void *smemDst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar, uint32_t bytes)
{
    const uint32_t b = smem_u32(bar), d = smem_u32(smemDst);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(d), "l"(map), "r"(x), "r"(y), "r"(z), "r"(b) : "memory");
}

// a further box on a barrier whose expected byte count already includes it
__device__ __forceinline__ void tma_copy_tile_3d(void *smemDst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(smemDst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

// Tensor maps reach the kernels as __grid_constant__ parameters, never through global memory: the TMA unit reads descriptors
// through a proxy of its own, and a descriptor REWRITTEN in place (the host does that whenever the frame size changes) can be
// served stale unless every consuming thread issues fence.proxy.tensormap::generic.acquire.sys first -- measured at +13 % on
// the whole step.  (Found by scripts/probe/soak_handle.py, seed 4031: a handle alternating between two frame sizes described a
// few pyramid levels of a few frames from boxes fetched with the other size's strides.)
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase)
{
    const uint32_t b = smem_u32(bar);
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra WAIT_%=;\n\t}"
                 ::"r"(b), "r"(phase) : "memory");
}

// ---- programmatic dependent launch: a kernel launched with the attribute below may start while its predecessor in the
// stream is still draining; it must not touch the predecessor's output before pdl_wait()
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ------------------------------------------------------------------------------------------
// pyramid resize, 11-bit fixed-point bilinear (A.1): level l from level l-1.
// Work unit = 128 x 32 output tile of one frame; its source region -- at most 192 x 48 bytes at scale
// factors up to 1.35 -- arrives by one TMA box load (cp.async.bulk.tensor, box start 16-byte aligned
// in x).  CTAs are persistent over a run of tiles (tile order: x tile, y tile, frame -- consecutive
// tiles of a run differ only in the frame, so the coefficient tables stay in registers): warp 4 is
// the producer, keeping a two-deep ring of boxes in flight (full/empty mbarriers), warps 0-3 consume.
// A consumer thread owns 4 adjacent output columns (one 32-bit store per output row) and 8 output
// rows, and walks the SOURCE rows of its band top to bottom: per source row it forms the four
// horizontal interpolations
//   h[x] = S[sx]*c0 + S[sx+1]*c1      (3 aligned word loads, 2 funnel shifts, 4 PRMT, 4 dp2a)
// once, keeps the previous row's in registers, and emits an output row whenever the pair
// (previous, current) is the pair (sy, sy+1) the next output row interpolates between -- every
// source row is loaded and interpolated exactly once per thread.  Downscaling makes sx and sy
// strictly increasing, so at most one output row is emitted per source row (checked on the host).
//   v = ((b0*(h0>>4))>>16) + ((b1*(h1>>4))>>16) + 2) >> 2, two pixels per register in 16-bit halves
// ------------------------------------------------------------------------------------------
#define RS_TW 128
#define RS_TH 32
#define RS_ROWS 8
#define RS_SRC_ROWS 12   // source rows a band of 8 output rows can touch at scale <= 1.35: 7*1.35 + 2
#define RS_BOXW 192
#define RS_BOXH 48
#define RS_STAGES 2

__device__ __forceinline__ void resize_hrow(const uint8_t *p, int sh, const uint32_t (&sel)[4], const uint32_t (&cc)[4], uint32_t (&hv)[4])
{
    const uint32_t w0 = *(const uint32_t *)p, w1 = *(const uint32_t *)(p + 4), w2 = *(const uint32_t *)(p + 8);
    const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);   // bytes sx .. sx+7
#pragma unroll
    for (int k = 0; k < 4; k++)
        hv[k] = __dp2a_lo(cc[k], __byte_perm(lo, hi, sel[k]), 0u) >> 4;                 // (c0*S[sx0] + c1*S[sx0+1]) >> 4
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(160)
k_resize(const __grid_constant__ CUtensorMap srcMap, int f0, uint8_t *__restrict__ pyr, long long slab, int dstOff, int dstPitch,
         int dw, int dh, const int4 *__restrict__ xtab, const int4 *__restrict__ ytab, int tilesY, int batch, int nTiles, int tilesPerCta)
{
    __shared__ __align__(128) uint8_t tileS[RS_STAGES][RS_BOXH * RS_BOXW];
    __shared__ __align__(8) uint64_t full[RS_STAGES], empty[RS_STAGES];
    __shared__ int4 sy[4][RS_ROWS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t0 = blockIdx.x * tilesPerCta, t1 = min(t0 + tilesPerCta, nTiles);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < RS_STAGES; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(smem_u32(&empty[s])) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_launch_dependents();       // the next level's kernel may set itself up; it waits for this grid before reading
    __syncthreads();

    if (warp == 4) {
        // ---- producer: one lane keeps RS_STAGES boxes in flight
        if (lane == 0) {
            int lastXY = -1, xs = 0, ys = 0;
            pdl_wait();            // the source level is complete and visible (only the TMA loads read it)
            for (int t = t0; t < t1; t++) {
                const int i = t - t0, b = i % RS_STAGES;
                const int xy = t / batch, f = t - xy * batch;
                if (xy != lastXY) {
                    const int txi = xy / tilesY, tyi = xy - txi * tilesY;
                    xs = __ldg(&xtab[txi * RS_TW]).x & ~15;      // box origin: first source column, 16-aligned
                    ys = __ldg(&ytab[tyi * RS_TH]).x;            //             first source row
                    lastXY = xy;
                }
                if (i >= RS_STAGES) mbar_wait(&empty[b], ((i / RS_STAGES) - 1) & 1);
                tma_load_tile_3d(tileS[b], &srcMap, xs, ys, f0 + f, &full[b], RS_BOXH * RS_BOXW);
            }
        }
        return;
    }

    // ---- consumers
    int lastXY = -1;
    int xs = 0, ys = 0, sxa = 0, sh = 0, x0 = 0, y0 = 0;
    uint32_t sel[4] = {0, 0, 0, 0}, cc[4] = {0, 0, 0, 0};
    for (int t = t0; t < t1; t++) {
        const int i = t - t0, b = i % RS_STAGES;
        const int xy = t / batch, f = t - xy * batch;
        if (xy != lastXY) {
            const int txi = xy / tilesY, tyi = xy - txi * tilesY;
            const int tx0 = txi * RS_TW, ty0 = tyi * RS_TH;
            x0 = tx0 + lane * 4; y0 = ty0 + warp * RS_ROWS;
            __syncwarp();
            if (lane < RS_ROWS) sy[warp][lane] = __ldg(&ytab[min(y0 + lane, dh - 1)]);   // {sy0, sy0+1, b0, b1}
            xs = __ldg(&xtab[tx0]).x & ~15;
            ys = __ldg(&ytab[ty0]).x;
            sxa = __ldg(&xtab[min(x0, dw - 1)]).x;                                       // this thread's first source column
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int4 tt = __ldg(&xtab[min(x0 + k, dw - 1)]);
                const uint32_t o = (uint32_t)(tt.x - sxa);                               // 0..4 at scale factors <= 1.35
                sel[k] = o | (o + 1) << 4;
                cc[k] = (uint32_t)tt.z;
            }
            sh = ((sxa - xs) & 3) * 8;
            lastXY = xy;
            __syncwarp();
        }
        mbar_wait(&full[b], (i / RS_STAGES) & 1);
        if (x0 < dw && y0 < dh) {
            uint8_t *dst = pyr + (size_t)f * slab + dstOff + (size_t)y0 * dstPitch + x0;
            const int yEnd = min(RS_ROWS, dh - y0);
            int4 ty = sy[warp][0];                                                       // uniform across the warp
            const int s0 = ty.x;
            const uint8_t *base = tileS[b] + (s0 - ys) * RS_BOXW + ((sxa - xs) & ~3);
            uint32_t prev[4], cur[4];
            resize_hrow(base, sh, sel, cc, prev);
            int r = 0;
#pragma unroll
            for (int j = 1; j < RS_SRC_ROWS; j++) {
                resize_hrow(base + j * RS_BOXW, sh, sel, cc, cur);
                if (ty.x - s0 == j - 1) {                                                // (prev, cur) = source rows (sy0, sy0+1) of output row r
                    // 16-bit pairs: (b0*h0 >> 16 | b0*h0' >> 16 << 16) + (b1*h1 ...) + (2 | 2 << 16), each half <= 1023
                    const uint32_t b0 = (uint32_t)ty.z, b1 = (uint32_t)ty.w;
                    const uint32_t w01 = __byte_perm(b0 * prev[0], b0 * prev[1], 0x7632) + __byte_perm(b1 * cur[0], b1 * cur[1], 0x7632) + 0x00020002u;
                    const uint32_t w23 = __byte_perm(b0 * prev[2], b0 * prev[3], 0x7632) + __byte_perm(b1 * cur[2], b1 * cur[3], 0x7632) + 0x00020002u;
                    *(uint32_t *)dst = __byte_perm(w01 >> 2, w23 >> 2, 0x6420);     // bytes past dw land in the pitch padding
                    dst += dstPitch;
                    if (++r >= yEnd) break;
                    ty = sy[warp][r];
                }
#pragma unroll
                for (int k = 0; k < 4; k++) prev[k] = cur[k];
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[b]);
    }
}

void launch_resize(const OrbxTensorMaps &srcMaps, int f0, uint8_t *pyr, const OrbxLayout &L, int level, const int4 *tabs, int batch, int nSM, cudaStream_t st)
{
    const OrbxLevel &d = L.lv[level];
    const int tilesX = (d.w + RS_TW - 1) / RS_TW, tilesY = (d.h + RS_TH - 1) / RS_TH;
    const int nTiles = tilesX * tilesY * batch;
    const int tilesPerCta = std::max(1, (nTiles + nSM * 8 - 1) / (nSM * 8));
    const int grid = (nTiles + tilesPerCta - 1) / tilesPerCta;
    launch_pdl(k_resize, dim3(grid), dim3(160), 0, st, srcMaps.m[level - 1], f0, pyr, L.slab, d.off, d.pitch, d.w, d.h,
               tabs + d.xtabOff, tabs + d.ytabOff, tilesY, batch, nTiles, tilesPerCta);
}

// ------------------------------------------------------------------------------------------
// 7x7 Gaussian blur, separable integer taps, REFLECT_101 (A.4):
//   out = min(255, (sum_y t[y] * (sum_x t[x] * I) + 32768) >> 16)
// One thread owns 4 adjacent columns (one 32-bit word of every output row) and walks BL_ROWS
// rows downwards.  Per input row it reads the 12 bytes x0-4 .. x0+7 as three aligned words,
// forms the four horizontal sums with ten dp4a against pre-shifted tap vectors (the data stays aligned), and
// keeps the last 7 of them per column in registers for the vertical sum -- the intermediate
// never touches shared or global memory.  A warp covers 128 columns; a CTA is 4 warps working
// on 4 vertically adjacent strips of one tile; one launch covers every level (tile table).
// ------------------------------------------------------------------------------------------
#define BL_ROWS ORBX_BLUR_ROWS   // 36: a band reads 36 + 6 = 42 = 6 x 7 input rows
struct BlurTaps { int t[7]; };

// Byte-permute selectors that apply REFLECT_101 at the right image edge to the 12-byte window
// {W0,W1,W2} = bytes x0-4 .. x0+7 (window index i = 0..11).  e = w - x0 is the number of valid
// bytes from x0 on; index i >= e+4 lies outside the row and equals index 2(e+3)-i.  Only three
// reflected bytes are ever consumed by valid outputs.  Interior threads get identity selectors,
// so every thread runs the same three PRMTs per row and no warp diverges at the border.
__device__ __forceinline__ void blur_edge_selectors(int e, uint32_t &sel1, uint32_t &selT, uint32_t &sel2)
{
    sel1 = 0x7654u; selT = 0x3210u; sel2 = 0x7654u;
    if (e >= 8) return;
    sel1 = 0; sel2 = 0; selT = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        int i = 4 + b;                                     // W1' byte b <- PRMT(W0, W1)
        int s = (i >= e + 4 && i <= e + 6) ? 2 * (e + 3) - i : i;
        sel1 |= (uint32_t)s << (4 * b);
        i = 8 + b;                                         // W2' byte b <- PRMT(T, W2), T = PRMT(W0, W1, selT)
        s = (i >= e + 4 && i <= e + 6) ? 2 * (e + 3) - i : i;
        if (s >= 8) sel2 |= (uint32_t)(4 + s - 8) << (4 * b);
        else { selT |= (uint32_t)s << (4 * b); sel2 |= (uint32_t)b << (4 * b); }
    }
}

// ------------------------------------------------------------------------------------------
// TMA staging (cp.async.bulk.tensor): one elected thread of the CTA fetches the whole tile --
// 128 columns + 2 x 16 bytes of halo, 4 x 36 rows + 6 rows of halo -- from the level's 3-D tensor map
// (x, y, frame) into shared memory; the hardware zero-fills what lies outside the image, and the
// REFLECT_101 rows/columns are resolved when the tile is READ (reflected rows are inside the box:
// the tile keeps 3 halo rows on both sides).  Completion is signalled on an mbarrier.
// ------------------------------------------------------------------------------------------
#define BL_BOXW 160   // bytes: 16 left + 128 + 16 right: TMA needs the box start 16-byte aligned in the inner dimension
#define BL_BOXH (4 * BL_ROWS + 6)   // rows: 3 + 4 bands + 3

// SAT: the taps sum to more than 256, so the result can exceed 255 and is saturated (cv::saturate_cast); with a sum of at most 256
// -- OpenCV 4's table -- the largest value is 32768 + 256 * 255 * 256 = 0xff8000: bits 16..23 are the result as they stand
template <bool SAT>
__global__ void __launch_bounds__(128)
k_blur(const __grid_constant__ OrbxTensorMaps tm, uint8_t *__restrict__ blur, const __grid_constant__ OrbxLayout L,
       const OrbxTile *__restrict__ tiles, BlurTaps taps, int f0)
{
    __shared__ __align__(128) uint8_t tileS[BL_BOXH * BL_BOXW];
    __shared__ __align__(8) uint64_t bar;
    const OrbxTile tile = tiles[blockIdx.x];
    const OrbxLevel &lv = L.lv[tile.level];
    const int f = f0 + blockIdx.y;
    const int w = lv.w, h = lv.h, pitch = lv.pitch;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        tma_load_tile_3d(tileS, &tm.m[tile.level], (int)tile.x0 * 4 - 16, (int)tile.y0 - 3, f, &bar, BL_BOXH * BL_BOXW);

    const int x0 = (tile.x0 + threadIdx.x) * 4;
    const int y0 = tile.y0 + threadIdx.y * BL_ROWS;
    const bool active = x0 < w && y0 < h;
    uint8_t *dst = blur + (size_t)f * L.slab + lv.off + x0;
    // horizontal taps as byte vectors against the ALIGNED words {W0,W1,W2} (window bytes 0..11): column x0+c takes window
    // bytes c+1 .. c+7, so byte b of word w carries tap 4w + b - (c+1) when that lies in 0..6 and zero otherwise -- the
    // data is never shifted, ten dp4a per four pixels and no byte permutes
    auto tapv = [&](int w, int c) {
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int k = 4 * w + b - (c + 1);
            if (k >= 0 && k <= 6) v |= (uint32_t)taps.t[k] << (8 * b);
        }
        return v;
    };
    const uint32_t TA0 = tapv(0, 0), TB0 = tapv(1, 0);
    const uint32_t TA1 = tapv(0, 1), TB1 = tapv(1, 1), TC1 = tapv(2, 1);
    const uint32_t TA2 = tapv(0, 2), TB2 = tapv(1, 2), TC2 = tapv(2, 2);
    const uint32_t TB3 = tapv(1, 3), TC3 = tapv(2, 3);
    const uint32_t T01 = (uint32_t)taps.t[0] | (uint32_t)taps.t[1] << 8, T23 = (uint32_t)taps.t[2] | (uint32_t)taps.t[3] << 8;
    const uint32_t T45 = (uint32_t)taps.t[4] | (uint32_t)taps.t[5] << 8, T6 = (uint32_t)taps.t[6];
    uint32_t sel1, selT, sel2;
    blur_edge_selectors(w - x0, sel1, selT, sel2);
    const bool leftEdge = x0 == 0;
    // warps that touch neither the left nor the right image edge skip the edge permutes altogether (uniform branch)
    const bool edgeWarp = __any_sync(0xffffffffu, active && (x0 == 0 || w - x0 < 8));
    const int rows = min(BL_ROWS, h - y0) + 6;
    // pp[c][s]: horizontal sums of (row before, row in slot s) packed as two 16-bit halves
    uint32_t pp[4][7], prev[4] = {0, 0, 0, 0};
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int k = 0; k < 7; k++) pp[c][k] = 0;
    mbar_wait(&bar, 0);
    if (!active) return;
    // shared row of image row g is (g - tile.y0 + 3); the row starts 16 bytes left of the tile, so this
    // thread's three words (bytes x0-4 .. x0+7) are words lane+3, lane+4, lane+5
    const uint32_t *ts = (const uint32_t *)tileS + threadIdx.x + 3;
    const int rowBias = 3 - (int)tile.y0;

    // one input row: horizontal sums of the thread's 4 columns from shared row p, vertical sum over the last
    // 7 rows; writes output row through dp when `store`
    auto row = [&](const int s, const uint32_t *p, const bool store, uint8_t *dp) {
        uint32_t W0 = p[0], W1 = p[1], W2 = p[2];
        if (edgeWarp) {
            if (leftEdge) W0 = __byte_perm(W1, W2, 0x1234);  // left edge: index -k equals index k
            const uint32_t T = __byte_perm(W0, W1, selT);
            W2 = __byte_perm(T, W2, sel2);
            W1 = __byte_perm(W0, W1, sel1);
        }
        // column x0+i needs bytes (i+1 .. i+7) of {W0,W1,W2}
        uint32_t hs[4];
        hs[0] = __dp4a(W1, TB0, __dp4a(W0, TA0, 0u));
        hs[1] = __dp4a(W2, TC1, __dp4a(W1, TB1, __dp4a(W0, TA1, 0u)));
        hs[2] = __dp4a(W2, TC2, __dp4a(W1, TB2, __dp4a(W0, TA2, 0u)));
        hs[3] = __dp4a(W2, TC3, __dp4a(W1, TB3, 0u));
        uint32_t acc[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            pp[c][s] = __byte_perm(prev[c], hs[c], 0x5410);
            prev[c] = hs[c];
            // rows r-6..r with taps 0..6: pairs end in slots of rows r-5, r-3, r-1; row r alone
            uint32_t a = 32768u + T6 * hs[c];
            a = __dp2a_lo(pp[c][(s + 2) % 7], T01, a);
            a = __dp2a_lo(pp[c][(s + 4) % 7], T23, a);
            a = __dp2a_lo(pp[c][(s + 6) % 7], T45, a);
            acc[c] = SAT ? min(a, 0x00ffffffu) : a;          // result byte = bits 16..23, saturated where it can overflow
        }
        if (store) {
            const uint32_t lo = __byte_perm(acc[0], acc[1], 0x0062), hi = __byte_perm(acc[2], acc[3], 0x0062);
            *(uint32_t *)dp = __byte_perm(lo, hi, 0x5410);
        }
    };

    // Input rows go in groups of 7 (the period of the register window).  A group whose rows all lie inside the image walks
    // straight pointers (interior bands read 42 = 6 x 7 rows that way); groups that touch the top or bottom edge
    // (REFLECT_101) or the end of a partial band take the general path.
#pragma unroll 1
    for (int r0 = 0; r0 < rows; r0 += 7) {
        const int g0 = y0 - 3 + r0;                                  // image row of the group's first input row
        if (g0 >= 0 && g0 + 6 < h && r0 + 7 <= rows) {
            const uint32_t *p = ts + (g0 + rowBias) * (BL_BOXW / 4);
            uint8_t *dp = dst + (size_t)(y0 + r0 - 6) * pitch;       // output row of input row r is y0 + r - 6
            if (r0 == 0) {
#pragma unroll
                for (int s = 0; s < 6; s++) row(s, p + s * (BL_BOXW / 4), false, dp);
                row(6, p + 6 * (BL_BOXW / 4), true, dp + (size_t)6 * pitch);
            } else {
#pragma unroll
                for (int s = 0; s < 7; s++) { row(s, p + s * (BL_BOXW / 4), true, dp); dp += pitch; }
            }
        } else {
#pragma unroll
            for (int s = 0; s < 7; s++) {
                const int r = r0 + s;
                if (r < rows) {
                    int g = y0 + r - 3;                              // REFLECT_101 (|overshoot| <= 3 < h)
                    g = g < 0 ? -g : (g >= h ? 2 * h - 2 - g : g);
                    row(s, ts + (g + rowBias) * (BL_BOXW / 4), r >= 6, dst + (size_t)(y0 + r - 6) * pitch);
                }
            }
        }
    }
}

void launch_blur(const OrbxTensorMaps &maps, uint8_t *blur, const OrbxLayout &L, const OrbxTile *tiles, int nTiles,
                 const int taps[7], int f0, int batch, cudaStream_t st)
{
    BlurTaps t;
    for (int k = 0; k < 7; k++) t.t[k] = taps[k];
    dim3 grid(nTiles, batch);
    int sum = 0;
    for (int k = 0; k < 7; k++) sum += taps[k];
    if (sum > 256) k_blur<true><<<grid, dim3(32, 4), 0, st>>>(maps, blur, L, tiles, t, f0);
    else k_blur<false><<<grid, dim3(32, 4), 0, st>>>(maps, blur, L, tiles, t, f0);
}

// ------------------------------------------------------------------------------------------
// Gridded FAST-9/16 with 3x3 NMS per cell and the ini/min threshold fallback (A.3).
// One CTA (128 threads) per (segment, frame); a segment is a run of horizontally adjacent cells of one cell
// row (<= 112 tested columns: 2-4 cells of 28-60 px), so the candidate lists of several cells share
// the CTA's lanes.  Output is not a keypoint list: DistributeOctTree in this fork only ever
// splits along y inside fixed x-strips (A.5), so all it needs per (strip, row) is the number of
// candidates and the best candidate (max response, first in the reference's emission order).
// Both are accumulated here with atomics:
//   cnt [frame][rowBase + strip*H + y]  += 1
//   best[frame][rowBase + strip*H + y]   = max(score<<56 | ~order<<28 | x<<14 | y)
// `order` = (cell index << 12 | yIn << 6 | xIn) is the position in the reference's emission
// order (cells row-major, raster inside a cell, orbextractor.cpp:930-968).
//
// Inside the CTA:
//   0  the run's window (144 bytes x hCell+6 rows) is fetched by one TMA box load (cp.async.bulk.tensor)
//   then twice, in the reference's order (orbextractor.cpp:940-957) -- every cell at the initial threshold, afterwards only the
//   cells whose NMS result came out empty at the minimum threshold:
//   1  quick reject on 4 pixels per item (packed bytes): a FAST-9 arc always contains ring
//      pixel k or k+8, so |I(p) - I(ring_k)| > t must hold for k in {0,8} and for k in {4,12};
//      survivors (5-14 % of the pixels at threshold 20) are appended to a list warp by warp
//   2  exact test and corner score in one pass: with d_k = I(p) - I(ring_k),
//        A = max( max_k min(d_k..d_k+8), -min_k max(d_k..d_k+8) )
//      is the largest threshold margin of the pixel: it is a corner at threshold t iff A > t, and its
//      cv::FAST response is A - 1.  A thread evaluates two survivors at once, one in each 16-bit half of its
//      registers, with three-input packed min/max (VIMNMX3.S16x2): 2 x (16 + 16) of them cover the 16 windows.
//   3  3x3 NMS on the run's score map (neighbours in another cell count as 0, as each cell is an
//      isolated cv::FAST call); a maximum goes straight into the row summaries
// ------------------------------------------------------------------------------------------
#define FS_T ORBX_FAST_THREADS  // threads per CTA
#define FW_P ORBX_FAST_PITCH  // shared window pitch in bytes = TMA box width (36 words: vertical neighbours are 4 banks apart)
#define FM_P ORBX_FAST_PITCH  // shared score-map pitch (the index arithmetic relies on FW_P == FM_P: list entries are yIn * FW_P + xs)
static_assert(FW_P == 144, "fast_row_of divides by 144");
__host__ __device__ __forceinline__ int fast_window_bytes(int winRows) { return (winRows * FW_P + 127) & ~127; }   // the score map behind it is a TMA destination too
__device__ __forceinline__ int fast_row_of(int e) { return (int)(((unsigned)e * 3641u) >> 19); }   // e / 144 for e < 9000 (60 rows)

__device__ __forceinline__ uint32_t swap16(uint32_t x) { return __byte_perm(x, x, 0x1032); }
__device__ __forceinline__ unsigned lanemask_lt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }
// shared-memory atomic add issued as-is (the compiler's own warp aggregation of atomicAdd is redundant where one lane adds for the warp)
__device__ __forceinline__ int smem_atomic_add(int *p, int v)
{
    int old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
    return old;
}

// largest threshold margins A of TWO pixels (window pitch FW_P), one per 16-bit half of every register; see stage 2 above.
// E[k] = (256 + vA - ringA_k) | (256 + vB - ringB_k) << 16 : both halves in [1, 511], so the packed subtraction never
// borrows across the halves and the packed three-input min/max (VIMNMX3.S16x2) serve both pixels at once.
__device__ __forceinline__ void fast_margin2(const uint8_t *cA, const uint8_t *cB, int &mA, int &mB)
{
    const uint32_t V2 = ((uint32_t)cA[0] | (uint32_t)cB[0] << 16) + 0x01000100u;
    uint32_t E[16];
#define ORBX_RING2(k, off) E[k] = V2 - ((uint32_t)cA[off] | (uint32_t)cB[off] << 16)
    ORBX_RING2(0, 3 * FW_P);       ORBX_RING2(1, 3 * FW_P + 1);   ORBX_RING2(2, 2 * FW_P + 2);   ORBX_RING2(3, FW_P + 3);
    ORBX_RING2(4, 3);              ORBX_RING2(5, -FW_P + 3);      ORBX_RING2(6, -2 * FW_P + 2);  ORBX_RING2(7, -3 * FW_P + 1);
    ORBX_RING2(8, -3 * FW_P);      ORBX_RING2(9, -3 * FW_P - 1);  ORBX_RING2(10, -2 * FW_P - 2); ORBX_RING2(11, -FW_P - 3);
    ORBX_RING2(12, -3);            ORBX_RING2(13, FW_P - 3);      ORBX_RING2(14, 2 * FW_P - 2);  ORBX_RING2(15, 3 * FW_P - 1);
#undef ORBX_RING2
    uint32_t tn[16], tx[16];
#pragma unroll
    for (int j = 0; j < 16; j++) {
        tn[j] = __vmins2(__vmins2(E[j], E[(j + 1) & 15]), E[(j + 2) & 15]);    // min / max of d_j..d_j+2
        tx[j] = __vmaxs2(__vmaxs2(E[j], E[(j + 1) & 15]), E[(j + 2) & 15]);
    }
    uint32_t a = 0u, b = 0x7fff7fffu;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        // window d_k..d_k+8
        a = __vmaxs2(a, __vmins2(__vmins2(tn[k], tn[(k + 3) & 15]), tn[(k + 6) & 15]));
        b = __vmins2(b, __vmaxs2(__vmaxs2(tx[k], tx[(k + 3) & 15]), tx[(k + 6) & 15]));
    }
    mA = max((int)(a & 0xffffu) - 256, 256 - (int)(b & 0xffffu));
    mB = max((int)(a >> 16) - 256, 256 - (int)(b >> 16));
}

// Stage 1 of k_fast_segs.  The items are (row, k-th aligned 4-pixel quad in play): all quads of the run in the first pass, the
// quads that touch a still-empty cell (qlist) in the second.  Thread t owns quad column k = t % nQx and the band of hB
// consecutive rows number t / nQx (R = FS_T / nQx bands, hB = ceil(hT / R); threads beyond R * nQx idle), so its border masks
// are constants and it walks DOWN its column: the centre words of rows y-3 .. y+3 stay in registers and a row costs three word
// loads (centre of row y+3, left and right neighbour word of row y).  Per item 4 VABSDIFF4 and bit 7 of ((d + K) | d) per
// byte, K = 127 - th: set iff d > th.  (A byte whose sum overflows carries one into its upper neighbour, which can only turn
// that neighbour's "d == th" into a pass -- the filter stays a superset; the overflowing byte itself has bit 7 of d set.)
// The survivor flags of up to 8 rows are collected in a register (4 bits per item) and written out warp by warp (order inside
// a warp: thread, row, pixel -- consecutive entries are vertical neighbours, FW_P / 4 = 36 words = 4 banks apart).  *ncand
// accumulates the number of survivors; the caller synchronises before reading the list.  Warp-synchronous: call with full warps.
__device__ __forceinline__ void fast_quick_reject(const uint8_t *win, uint16_t *cand, int *ncand, const uint8_t *qlist, const bool dense,
                                                 const int nQx, const unsigned mQx, const int hB, const int tid, const int lane,
                                                 const int B0, const int wT, const int hT, const int th)
{
    // quads are aligned to shared-memory words: the first and last quad of a row may be partly outside
    const int wq0 = B0 >> 2, nQ = ((B0 + wT - 1) >> 2) - wq0 + 1;
    const int band = (int)(((unsigned)tid * mQx) >> 20), k = tid - band * nQx;     // mQx = 2^20 / nQx + 1
    const int R = (int)(((unsigned)FS_T * mQx) >> 20);
    const int q = dense ? k : (int)qlist[k];
    uint32_t mask = 0x80808080u;
    if (q == 0) mask &= ~((1u << (8 * (B0 & 3))) - 1u);
    if (q == nQ - 1) {
        const int nLast = ((B0 + wT - 1) & 3) + 1;
        if (nLast < 4) mask &= (1u << (8 * nLast)) - 1u;
    }
    const uint32_t K = (uint32_t)(127 - th) * 0x01010101u;
    constexpr int P4 = FW_P / 4;
    for (int rb = 0; rb < hB; rb += 8) {
        const int y0 = band * hB + rb;                                  // first row of this thread in this round
        const int n = band < R ? min(min(8, hB - rb), hT - y0) : 0;
        uint32_t bits = 0;
        if (n > 0) {
            const uint32_t *rw = (const uint32_t *)(win + (y0 + 3) * FW_P) + wq0 + q;
            uint32_t cm3 = rw[-3 * P4], cm2 = rw[-2 * P4], cm1 = rw[-P4], c0 = rw[0], cp1 = rw[P4], cp2 = rw[2 * P4];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (j >= n) break;
                const uint32_t cp3 = rw[(j + 3) * P4];
                const uint32_t W0 = rw[j * P4 - 1], W2 = rw[j * P4 + 1];
                const uint32_t d0 = __vabsdiffu4(c0, cp3), d8 = __vabsdiffu4(c0, cm3);
                const uint32_t d4 = __vabsdiffu4(c0, __byte_perm(c0, W2, 0x6543)), d12 = __vabsdiffu4(c0, __byte_perm(W0, c0, 0x4321));
                const uint32_t X = (d0 + K) | d0 | (d8 + K) | d8;
                const uint32_t Y = (d4 + K) | d4 | (d12 + K) | d12;
                const uint32_t pass = X & Y & mask;
                // flag bits 7,15,23,31 -> one nibble: the products land on distinct bits, the top four are the flags
                bits |= ((pass * 0x00204081u) >> 28) << (4 * j);
                cm3 = cm2; cm2 = cm1; cm1 = c0; c0 = cp1; cp1 = cp2; cp2 = cp3;
            }
        }
        // the warp's survivors go to the list as one block: inclusive scan of the per-thread counts, ONE shared atomic per warp
        // for the block's position (the order of the warps' blocks does not reach the results: scores go to the score map, the
        // row summaries are order-independent), no block barrier
        const int c = __popc(bits);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        int base = 0;
        if (lane == 31 && incl) base = smem_atomic_add(ncand, incl);
        base = __shfl_sync(0xffffffffu, base, 31);
        int pos = base + incl - c;
        const int e0 = y0 * FW_P + 4 * (wq0 + q) - B0;                 // entry yIn * FW_P + xs of pixel 0 of the thread's first row
        while (bits) {                                                 // two survivors per trip
            const int b = __ffs((int)bits) - 1;
            bits &= bits - 1u;
            cand[pos] = (uint16_t)(e0 + (b >> 2) * FW_P + (b & 3));
            if (bits) {
                const int b2 = __ffs((int)bits) - 1;
                bits &= bits - 1u;
                cand[pos + 1] = (uint16_t)(e0 + (b2 >> 2) * FW_P + (b2 & 3));
            }
            pos += 2;
        }
    }
}

// a hint of nine blocks gives ptxas 48 registers to schedule with (45 with a hint of ten); ten CTAs of 128 threads still fit an SM
__global__ void __launch_bounds__(FS_T, 9)
k_fast_segs(const __grid_constant__ OrbxTensorMaps tm, int f0, const __grid_constant__ OrbxLayout L,
            const OrbxSeg *__restrict__ segs, uint32_t *__restrict__ cnt,
            unsigned long long *__restrict__ best, OrbxDbgCand *__restrict__ dbg,
            int *__restrict__ dbgCount, int dbgCap, int winRows, int listCap, int kcap)
{
    extern __shared__ __align__(128) uint8_t fsm[];
    uint8_t *win = fsm;                                              // winRows x FW_P (TMA destination)
    uint8_t *smap = win + fast_window_bytes(winRows);                // (winRows - 4) x FM_P: scores with a zero border (128-byte aligned: TMA writes it)
    uint16_t *cand = (uint16_t *)(smap + (winRows - 4) * FM_P);      // listCap: yIn * FW_P + xs
    uint16_t *corner = cand + listCap;                               // kcap (<= listCap): on overflow stage 3a scans the score map instead
    __shared__ __align__(8) uint64_t bar;
    __shared__ int ncand, ncorner;                                    // stage 1 / stage 2 list lengths
    __shared__ int nq2s;
    __shared__ unsigned cellsDone;
    __shared__ uint8_t cellOf[ORBX_SEG_W], qlist[64];

    const OrbxSeg seg = segs[blockIdx.x];
    const int frame = blockIdx.y;
    const OrbxLevel &lv = L.lv[seg.level];
    const int wT = seg.wT, hT = seg.hT;
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned lt = lanemask_lt();
    const int wCell = lv.wCell;

    // ---- stage 0: TMA fetches the window from the level's tensor map.  The box must start 16-byte aligned
    // in x and the word left of the first tested pixel's word is read too: box x = (x0 - 4) & ~15, and tested
    // column xs sits at shared byte B0 + xs of its row.
    const int bx = ((int)seg.x0 - 4) & ~15, B0 = (int)seg.x0 + 3 - bx;
    // Only thread 0 touches the mbarrier -- it initialises it, issues the load and, after its share of the set-up work, waits for
    // the bytes; the block barrier behind that hands the window to everybody else (one barrier instead of two).
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        ncand = 0; ncorner = 0; cellsDone = 0u;
        // the score map is zeroed by the TMA unit as well: a second box of the same shape from far outside the level is all
        // out-of-bounds fill (1.5 % faster than 128 threads storing zeros).  It is winH = hT + 6 rows, four more than the
        // score map: the surplus lands in the front of the survivor list, which nobody writes before the barrier below.
        tma_load_tile_3d(win, &tm.m[seg.level], bx >> 2, (int)seg.y0, f0 + frame, &bar, 2 * FW_P * lv.winH);   // x in 32-bit elements
        tma_copy_tile_3d(smap, &tm.m[seg.level], bx >> 2, -4096, f0 + frame, &bar);
    }
    if (tid < wT) cellOf[tid] = (uint8_t)(((unsigned)tid * (unsigned)lv.cellMagic) >> 16);            // tid / wCell
    const int nCells = (int)(((unsigned)(wT - 1) * (unsigned)lv.cellMagic) >> 16) + 1;
    const unsigned allCells = (1u << nCells) - 1u;
    if (tid == 0) mbar_wait(&bar, 0);
    __syncthreads();

    uint32_t *cntF = cnt + (size_t)frame * L.rowsPerFrame + lv.rowBase;
    unsigned long long *bestF = best + (size_t)frame * L.rowsPerFrame + lv.rowBase;
    const int nQ = ((B0 + wT - 1) >> 2) - (B0 >> 2) + 1;

    // The reference runs cv::FAST per cell at the initial threshold and only where that leaves nothing (after NMS) again at
    // the minimum threshold (orbextractor.cpp:940-957).  Same order here: pass 0 = all cells at iniTh, pass 1 = the cells
    // of the run that came out empty, at minTh.  (A cell whose ini-corners all fell to NMS ties is such a cell: pass 1
    // recomputes its corners above minTh from scratch, the ones above iniTh included.)
    unsigned cmask = allCells;
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
        const int th = pass ? L.minTh : L.iniTh;
        const int thQ = min(th, 127);                                 // the packed quick reject compares 7-bit fields; a lower
                                                                      // threshold only lets more pixels through to the exact test
        bool dense = true;
        unsigned mQx = seg.mQ & 0xffffffu;
        int nQx = nQ, hB = (int)(seg.mQ >> 24);
        if (pass) {
            cmask = allCells & ~cellsDone;                            // final since the barrier that ends pass 0
            if (!cmask) break;                                        // uniform
            dense = cmask == allCells;
            if (tid == 0) { ncand = 0; ncorner = 0; }
            if (!dense && tid < 32) {
                // aligned quads that touch an empty cell, in ascending order (two per lane)
                const int wq0 = B0 >> 2;
                bool in0 = false, in1 = false;
                if (lane < nQ) {
                    const int a = max(4 * (wq0 + lane) - B0, 0), b = min(4 * (wq0 + lane) + 3 - B0, wT - 1);
                    in0 = ((cmask >> cellOf[a]) | (cmask >> cellOf[b])) & 1u;
                }
                if (lane + 32 < nQ) {
                    const int a = max(4 * (wq0 + lane + 32) - B0, 0), b = min(4 * (wq0 + lane + 32) + 3 - B0, wT - 1);
                    in1 = ((cmask >> cellOf[a]) | (cmask >> cellOf[b])) & 1u;
                }
                const unsigned b0 = __ballot_sync(0xffffffffu, in0), b1 = __ballot_sync(0xffffffffu, in1);
                if (in0) qlist[__popc(b0 & lt)] = (uint8_t)lane;
                if (in1) qlist[__popc(b0) + __popc(b1 & lt)] = (uint8_t)(lane + 32);
                if (lane == 0) nq2s = __popc(b0) + __popc(b1);
            }
            __syncthreads();
            if (!dense) {
                nQx = nq2s; mQx = (1u << 20) / (unsigned)nQx + 1u;
                const int R = FS_T / nQx;
                hB = (hT + R - 1) / R;
            }
        }

        // ---- stage 1: packed quick reject, 4 pixels per item
        fast_quick_reject(win, cand, &ncand, qlist, dense, nQx, mQx, hB, tid, lane, B0, wT, hT, thQ);
        __syncthreads();
        const int nc = ncand;

        // ---- stage 2: threshold margin of every survivor, two per thread (survivors i and i + half share the
        // 16-bit halves of the registers); corners (margin > threshold, in a cell of this pass) get their score written
        // to the score map and are compacted (one shared atomic per warp)
        const int half = (nc + 1) >> 1;
        for (int i0 = 0; i0 < half; i0 += FS_T) {
            const int i = i0 + tid;
            int eA = 0, eB = 0, mA = 0, mB = 0;
            bool hasB = false;
            if (i < half) {
                eA = cand[i];
                hasB = i + half < nc;
                eB = hasB ? cand[i + half] : eA;
                // e = yIn * FW_P + xs and both pitches are equal: window byte = e + 3 rows + B0, score-map byte = e + 1 row + 1
                fast_margin2(win + eA + (3 * FW_P + B0), win + eB + (3 * FW_P + B0), mA, mB);
            }
            bool cornerA = mA > th, cornerB = hasB && mB > th;
            if (pass) {                                               // a quad may reach into a neighbouring cell that is done
                cornerA = cornerA && ((cmask >> cellOf[eA - fast_row_of(eA) * FW_P]) & 1u);
                cornerB = cornerB && ((cmask >> cellOf[eB - fast_row_of(eB) * FW_P]) & 1u);
            }
            const unsigned balA = __ballot_sync(0xffffffffu, cornerA), balB = __ballot_sync(0xffffffffu, cornerB);
            if (balA | balB) {
                int base = 0;
                if (lane == 0) base = smem_atomic_add(&ncorner, __popc(balA) + __popc(balB));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (cornerA) {
                    const int at = base + __popc(balA & lt);
                    if (at < kcap) corner[at] = (uint16_t)eA;
                    smap[eA + (FM_P + 1)] = (uint8_t)(mA - 1);
                }
                if (cornerB) {
                    const int at = base + __popc(balA) + __popc(balB & lt);
                    if (at < kcap) corner[at] = (uint16_t)eB;
                    smap[eB + (FM_P + 1)] = (uint8_t)(mB - 1);
                }
            }
        }
        __syncthreads();
        const int nk = ncorner;

        // ---- stage 3: 3x3 non-maximum suppression (strict >; outside the cell interior counts as 0); a maximum goes straight
        // into the per-(strip,row) summaries -- its cell is final in this pass: pass 0 keeps every cell that has one, pass 1
        // only works on cells that had none.
        // (a corner list that did not fit -- more than kcap corners, dense noise -- is replaced by a scan of the score map)
        const bool scan = nk > kcap;
        const int nIt = scan ? wT * hT : nk;
        unsigned done = 0u;
        for (int i0 = 0; i0 < nIt; i0 += FS_T) {
            const int i = i0 + tid;
            bool have = i < nIt;
            int e = 0, xs = 0, yIn = 0;
            if (have) {
                if (!scan) { e = corner[i]; yIn = fast_row_of(e); xs = e - yIn * FW_P; }
                else {
                    yIn = i / wT;
                    xs = i - yIn * wT; e = yIn * FW_P + xs;
                    have = smap[e + (FM_P + 1)] != 0 && ((cmask >> cellOf[xs]) & 1u);
                }
            }
            if (have) {
                const int cl = cellOf[xs], xIn = xs - cl * wCell;
                const uint8_t *sp = smap + e + (FM_P + 1);
                const int v = sp[0];
                const bool lOk = xIn > 0, rOk = xIn < wCell - 1;
                const int l0 = lOk ? max(max((int)sp[-FM_P - 1], (int)sp[-1]), (int)sp[FM_P - 1]) : 0;
                const int r0 = rOk ? max(max((int)sp[-FM_P + 1], (int)sp[1]), (int)sp[FM_P + 1]) : 0;
                const int m = max(max(l0, r0), max((int)sp[-FM_P], (int)sp[FM_P]));
                if (v > m) {
                    done |= 1u << cl;                                  // the cell keeps its initial-threshold result
                    const int xr = seg.cj0 * wCell + 3 + xs, yr = seg.ci * lv.hCell + 3 + yIn;     // relative to (16,16), :963-964
                    int strip = 0;                                                                  // xr / hX, :710
#pragma unroll
                    for (int k = 1; k < ORBX_MAX_STRIPS; k++) strip += (xr >= k * lv.hX);
                    const int row = strip * lv.H + yr;
                    const unsigned order = (unsigned)(seg.ci * lv.nCols + seg.cj0 + cl) << 12 | (unsigned)(yIn << 6 | xIn);
                    const unsigned long long key = ((unsigned long long)v << 56) |
                                                   ((unsigned long long)(0x0fffffffu - order) << 28) |
                                                   ((unsigned long long)xr << 14) | (unsigned long long)yr;
                    atomicAdd(&cntF[row], 1u);
                    atomicMax(&bestF[row], key);
                    if (dbg) {
                        const int slot = frame * L.nlevels + seg.level;
                        const int pos = atomicAdd(&dbgCount[slot], 1);
                        if (pos < dbgCap) {
                            OrbxDbgCand c; c.xy = xr | (yr << 16); c.score = v;
                            dbg[(size_t)slot * dbgCap + pos] = c;
                        }
                    }
                }
            }
        }
        if (pass == 0) {
            done = __reduce_or_sync(0xffffffffu, done);
            if (lane == 0 && done) atomicOr(&cellsDone, done);
            __syncthreads();                                          // cellsDone is complete before pass 1 reads it
        }
    }
}

// Capacity of the corner list.  The survivor list must hold every tested pixel of a run (noise lets them all through the quick
// reject); the corner list gets what is left of the shared memory that still admits ten (else nine, ...) resident CTAs of 128
// threads per SM, but at least a sixth of the pixels -- past that stage 3 falls back to scanning the score map.  ORBX_FAST_KCAP overrides (tests).
int fast_corner_cap(int winRows, int listCap)
{
    static const int forced = getenv("ORBX_FAST_KCAP") ? atoi(getenv("ORBX_FAST_KCAP")) : 0;
    if (forced > 0) return std::min(listCap, (forced + 7) & ~7);
    const long fixed = (long)fast_window_bytes(winRows) + (long)(winRows - 4) * FM_P + 2L * listCap;
    for (int ctas = 10; ctas >= 6; ctas--) {
        const long budget = (233472 - ctas * 1024) / ctas - 512;   // per CTA: 228 KB per SM, 1 KB reserved per CTA, static shared memory
        const long room = (budget - fixed) / 2;
        if (room >= listCap / 6) return (int)std::min<long>(listCap, room & ~7L);
    }
    return listCap;
}

size_t fast_smem_bytes(int winRows, int listCap)
{
    const size_t lists = (size_t)listCap * 2 + (size_t)fast_corner_cap(winRows, listCap) * 2;
    // the zero box written over the score map is a full window (four rows more than the map): tiny levels keep room for it
    return (size_t)fast_window_bytes(winRows) + std::max((size_t)(winRows - 4) * FM_P + lists, (size_t)winRows * FW_P);
}

cudaError_t launch_fast(const OrbxTensorMaps &maps, int f0, const OrbxLayout &L, const OrbxSeg *segs, int segBegin, int segCount,
                        uint32_t *cnt, unsigned long long *best, OrbxDbgCand *dbg, int *dbgCount, int dbgCap,
                        int winRows, int listCap, int batch, cudaStream_t st)
{
    if (segCount <= 0) return cudaSuccess;
    const size_t smem = fast_smem_bytes(winRows, listCap);
    {
        cudaError_t e = orbx_raise_dyn_smem((const void *)k_fast_segs, smem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid(segCount, batch);
    k_fast_segs<<<grid, FS_T, smem, st>>>(maps, f0, L, segs + segBegin, cnt, best, dbg, dbgCount, dbgCap, winRows, listCap,
                                          fast_corner_cap(winRows, listCap));
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// DistributeOctTree (A.5): one CTA per (level, frame); warp 0 replays the reference's list
// algorithm round by round with warp-synchronous scans (no block barriers inside the rounds), the
// other warps only help with the row prefix sums in front and with the output stage behind.
// A node is (strip, [y0,y1)); its size is a difference of the per-strip prefix sums P of the row
// counts, so dividing a node is O(1).
//   main pass   : every node with >1 point is divided; children go to the list front in the
//                 order n2, n4 (=> n4 first), parents processed front to back     (:735-808)
//   priority    : nodes ordered by (size, creation) descending, divided until size >= N (:814-878).
//                 The order comes from a stable LSD counting sort on the size (10 bits per pass,
//                 __match_any ranks inside a chunk of 32), started from the tie order.
//   output      : per node the max-response point, first in emission order on ties  (:885-901)
// Under a monotone allocator "higher address" == "created later"; all expandable nodes of a
// round were created in the previous round and sit at the list front in reverse creation order,
// so "newest first" == "lowest list position first".
// ------------------------------------------------------------------------------------------
#define OCT_T 128
#define OCT_BUCKETS 1024
// bytes of P, the sort buckets, the two node lists, the two order arrays and the flags, rounded up to 16
__host__ __device__ inline size_t octree_lists_bytes(int maxRows, int maxNodes)
{
    return ((size_t)((maxRows + 4) & ~3) * 4 + (size_t)OCT_BUCKETS * 4 + (size_t)maxNodes * (4 * 2 + 2 * 2 + 1) + 15) & ~(size_t)15;
}
__device__ __forceinline__ uint32_t node_pack(int s, int y0, int y1) { return (uint32_t)s << 26 | (uint32_t)y0 << 13 | (uint32_t)y1; }

struct OctSh {
    int n, which;
    int scan[34];
};

// (the minimum-blocks hint widens ptxas' register budget: 77 registers instead of 64, the sequential rounds of warp 0 run 4 % faster;
// the same hint makes k_describe and k_resize slower -- 127 / 72 registers -- and leaves k_blur unchanged)
__global__ void __launch_bounds__(OCT_T, 4)
k_octree(const __grid_constant__ OrbxLayout L, const uint32_t *__restrict__ cnt,
         const unsigned long long *__restrict__ best, int2 *__restrict__ slots,
         int *__restrict__ lvlCount, int maxRows, int maxNodes)
{
    extern __shared__ __align__(16) unsigned char sm_raw[];
    __shared__ OctSh sh;
    const int level = blockIdx.x, frame = blockIdx.y;
    const OrbxLevel &lv = L.lv[level];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = lv.H, nR = lv.nIni * H, N = lv.quota;
    const unsigned lt = lanemask_lt();

    int *P = (int *)sm_raw;                                          // maxRows + 1 (padded to 4)
    int *bucket = P + ((maxRows + 4) & ~3);                          // OCT_BUCKETS
    uint32_t *listA = (uint32_t *)(bucket + OCT_BUCKETS);            // maxNodes
    uint32_t *listB = listA + maxNodes;
    uint16_t *ordA = (uint16_t *)(listB + maxNodes);                 // maxNodes each
    uint16_t *ordB = ordA + maxNodes;
    uint8_t *done = (uint8_t *)(ordB + maxNodes);                    // maxNodes
    unsigned long long *bestS = (unsigned long long *)(sm_raw + octree_lists_bytes(maxRows, maxNodes));   // maxRows

    const uint32_t *cntF = cnt + (size_t)frame * L.rowsPerFrame + lv.rowBase;
    const unsigned long long *bestF = best + (size_t)frame * L.rowsPerFrame + lv.rowBase;

    for (int i = tid; i < nR; i += OCT_T) P[i] = (int)cntF[i];
    __syncthreads();
    const int totalCand = block_excl_scan(P, nR, sh.scan);
    if (tid == 0) P[nR] = totalCand;
    __syncthreads();

#define NODE_DECODE(node) const int s_ = (node) >> 26, y0_ = ((node) >> 13) & 8191, y1_ = (node)&8191; const int pb_ = s_ * H
#define NODE_SIZE() (P[pb_ + y1_] - P[pb_ + y0_])

    if (warp == 0) {
        uint32_t *cur = listA, *nxt = listB;
        int n = 0;
        {   // :691-726 initial strips, empty ones erased
            const bool has = lane < lv.nIni && P[(lane + 1) * H] - P[lane * H] > 0;
            const unsigned b = __ballot_sync(0xffffffffu, has);
            if (has) cur[__popc(b & lt)] = node_pack(lane, 0, H);
            n = __popc(b);
        }
        __syncwarp();
        bool finish = (n == 0);
        while (!finish) {
            const int prevSize = n;
            // ---------------- main pass, sweep 1: totals
            int totC = 0, totN = 0, rec = 0;
            for (int i = lane; i < n; i += 32) {
                const uint32_t node = cur[i];
                NODE_DECODE(node);
                const int sz = NODE_SIZE();
                if (sz > 1) {
                    const int mid = y0_ + ((y1_ - y0_) >> 1);      // :75 integer halfY
                    const int c2 = P[pb_ + mid] - P[pb_ + y0_], c4 = sz - c2;
                    totC += (c2 > 0) + (c4 > 0);
                    rec += (c2 > 1) + (c4 > 1);
                } else totN++;
            }
            totC = __reduce_add_sync(0xffffffffu, totC);
            totN = __reduce_add_sync(0xffffffffu, totN);
            rec = __reduce_add_sync(0xffffffffu, rec);
            // ---------------- sweep 2: children to the front (later parents nearer the front), untouched nodes behind
            int runC = 0, runN = 0;
            for (int i0 = 0; i0 < n; i0 += 32) {
                const int i = i0 + lane;
                int nch = 0, c2 = 0, c4 = 0, mid = 0;
                bool leaf = false;
                uint32_t node = 0;
                if (i < n) {
                    node = cur[i];
                    NODE_DECODE(node);
                    const int sz = NODE_SIZE();
                    if (sz > 1) {
                        mid = y0_ + ((y1_ - y0_) >> 1);
                        c2 = P[pb_ + mid] - P[pb_ + y0_]; c4 = sz - c2;
                        nch = (c2 > 0) + (c4 > 0);
                    } else leaf = true;
                }
                const unsigned b1 = __ballot_sync(0xffffffffu, nch == 1), b2 = __ballot_sync(0xffffffffu, nch == 2);
                const unsigned bl = __ballot_sync(0xffffffffu, leaf);
                if (nch) {
                    NODE_DECODE(node); (void)pb_;
                    int pos = totC - (runC + __popc(b1 & lt) + 2 * __popc(b2 & lt) + nch);
                    if (c4 > 0) nxt[pos++] = node_pack(s_, mid, y1_);
                    if (c2 > 0) nxt[pos] = node_pack(s_, y0_, mid);
                } else if (leaf) {
                    nxt[totC + runN + __popc(bl & lt)] = node;
                }
                runC += __popc(b1) + 2 * __popc(b2);
                runN += __popc(bl);
            }
            __syncwarp();
            { uint32_t *t = cur; cur = nxt; nxt = t; }
            n = totC + totN;
            const int nToExpand = rec;
            if (n > maxNodes) n = maxNodes; // cannot happen for supported shapes (see orbx_api geometry checks)

            if (n >= N || n == prevSize) {
                finish = true;
            } else if (n + nToExpand * 3 > N) {
                // ---------------- priority rounds
                while (!finish) {
                    const int prev2 = n;
                    // expandable nodes in list order, largest size
                    int R = 0, maxSz = 0;
                    for (int i0 = 0; i0 < n; i0 += 32) {
                        const int i = i0 + lane;
                        int sz = 0;
                        if (i < n) { const uint32_t node = cur[i]; NODE_DECODE(node); sz = NODE_SIZE(); done[i] = 0; }
                        const unsigned b = __ballot_sync(0xffffffffu, sz > 1);
                        if (sz > 1) ordA[R + __popc(b & lt)] = (uint16_t)i;
                        R += __popc(b);
                        maxSz = max(maxSz, sz);
                    }
                    if (R == 0) { finish = true; break; }   // nothing expandable: size unchanged (:875)
                    maxSz = __reduce_max_sync(0xffffffffu, maxSz);
                    __syncwarp();
                    // tie order: lowest list position first (tieRule 0) or highest first (tieRule 1)
                    if (L.tieRule) {
                        for (int k = lane; k < R / 2; k += 32) { const uint16_t a = ordA[k], b = ordA[R - 1 - k]; ordA[k] = b; ordA[R - 1 - k] = a; }
                        __syncwarp();
                    }
                    // stable counting sort by size, descending, 10 bits per pass
                    uint16_t *src = ordA, *dst = ordB;
                    for (int shift = 0; (maxSz >> shift) > 0; shift += 10) {
                        for (int k = lane; k < OCT_BUCKETS / 4; k += 32) ((int4 *)bucket)[k] = make_int4(0, 0, 0, 0);
                        __syncwarp();
                        for (int k = lane; k < R; k += 32) {
                            const uint32_t node = cur[src[k]];
                            NODE_DECODE(node);
                            atomicAdd(&bucket[(OCT_BUCKETS - 1) - ((NODE_SIZE() >> shift) & (OCT_BUCKETS - 1))], 1);
                        }
                        __syncwarp();
                        {   // exclusive scan of the buckets: lane owns 32 consecutive ones
                            int4 v[8];
                            int sum = 0;
#pragma unroll
                            for (int q = 0; q < 8; q++) { v[q] = ((int4 *)bucket)[lane * 8 + q]; sum += v[q].x + v[q].y + v[q].z + v[q].w; }
                            int incl = sum;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
                            int run = incl - sum;
#pragma unroll
                            for (int q = 0; q < 8; q++) {
                                int4 w;
                                w.x = run; run += v[q].x; w.y = run; run += v[q].y; w.z = run; run += v[q].z; w.w = run; run += v[q].w;
                                ((int4 *)bucket)[lane * 8 + q] = w;
                            }
                        }
                        __syncwarp();
                        for (int k0 = 0; k0 < R; k0 += 32) {
                            const int k = k0 + lane;
                            const unsigned act = __ballot_sync(0xffffffffu, k < R);
                            if (k < R) {
                                const int i = src[k];
                                const uint32_t node = cur[i];
                                NODE_DECODE(node);
                                const int d = (OCT_BUCKETS - 1) - ((NODE_SIZE() >> shift) & (OCT_BUCKETS - 1));
                                const unsigned peers = __match_any_sync(act, d);
                                const int rank = __popc(peers & lt);
                                const int base = bucket[d];
                                __syncwarp(act);
                                if (rank == 0) bucket[d] = base + __popc(peers);
                                dst[base + rank] = (uint16_t)i;
                            }
                            __syncwarp();
                        }
                        { uint16_t *t = src; src = dst; dst = t; }
                    }
                    // sorted position j -> children count; stop once the list would reach N (:871)
                    int jstar = R - 1, totC2 = 0, runG = 0;
                    for (int j0 = 0; j0 < R; j0 += 32) {
                        const int j = j0 + lane;
                        int nch = 0;
                        if (j < R) {
                            const uint32_t node = cur[src[j]];
                            NODE_DECODE(node);
                            const int sz = NODE_SIZE();
                            const int mid = y0_ + ((y1_ - y0_) >> 1);
                            const int c2 = P[pb_ + mid] - P[pb_ + y0_], c4 = sz - c2;
                            nch = (c2 > 0) + (c4 > 0);
                        }
                        const unsigned b1 = __ballot_sync(0xffffffffu, nch == 1), b2 = __ballot_sync(0xffffffffu, nch == 2);
                        const unsigned le = lt | (1u << lane);
                        const bool hit = j < R && n + runG + __popc(b2 & le) >= N;      // gain = nch - 1
                        const unsigned bh = __ballot_sync(0xffffffffu, hit);
                        if (bh) {
                            const int l = __ffs(bh) - 1;
                            const unsigned upto = (2u << l) - 1u;
                            jstar = j0 + l;
                            totC2 += __popc(b1 & upto) + 2 * __popc(b2 & upto);
                            break;
                        }
                        runG += __popc(b2);
                        totC2 += __popc(b1) + 2 * __popc(b2);
                    }
                    // children of the divided nodes: later-processed parents' children sit nearer the front
                    int runC2 = 0;
                    for (int j0 = 0; j0 <= jstar; j0 += 32) {
                        const int j = j0 + lane;
                        int nch = 0, c2 = 0, c4 = 0, mid = 0;
                        uint32_t node = 0;
                        if (j <= jstar) {
                            const int i = src[j];
                            done[i] = 1;
                            node = cur[i];
                            NODE_DECODE(node);
                            const int sz = NODE_SIZE();
                            mid = y0_ + ((y1_ - y0_) >> 1);
                            c2 = P[pb_ + mid] - P[pb_ + y0_]; c4 = sz - c2;
                            nch = (c2 > 0) + (c4 > 0);
                        }
                        const unsigned b1 = __ballot_sync(0xffffffffu, nch == 1), b2 = __ballot_sync(0xffffffffu, nch == 2);
                        if (nch) {
                            NODE_DECODE(node); (void)pb_;
                            int pos = totC2 - (runC2 + __popc(b1 & lt) + 2 * __popc(b2 & lt) + nch);
                            if (c4 > 0) nxt[pos++] = node_pack(s_, mid, y1_);
                            if (c2 > 0) nxt[pos] = node_pack(s_, y0_, mid);
                        }
                        runC2 += __popc(b1) + 2 * __popc(b2);
                    }
                    __syncwarp();
                    // the rest keeps its list order behind the new children
                    int runU = 0;
                    for (int i0 = 0; i0 < n; i0 += 32) {
                        const int i = i0 + lane;
                        const bool keep = i < n && !done[i];
                        const unsigned b = __ballot_sync(0xffffffffu, keep);
                        if (keep) nxt[totC2 + runU + __popc(b & lt)] = cur[i];
                        runU += __popc(b);
                    }
                    __syncwarp();
                    { uint32_t *t = cur; cur = nxt; nxt = t; }
                    n = totC2 + runU;
                    if (n > maxNodes) n = maxNodes;
                    if (n >= N || n == prev2) finish = true;
                }
            }
        }
        if (lane == 0) { sh.n = n; sh.which = (cur == listA) ? 0 : 1; }
    } else {
        // meanwhile the other warps bring the per-row best candidates into shared memory for the output stage
        for (int i = tid - 32; i < nR; i += OCT_T - 32) bestS[i] = __ldg(&bestF[i]);
    }
    __syncthreads();

    // ---------------- output: best point of each node, list order (:885-901); 8 lanes per node
    {
        const int n = sh.n;
        const uint32_t *cur = sh.which ? listB : listA;
        int2 *out = slots + (size_t)frame * L.slotsPerFrame + lv.slotBase;
        const int nOut = n < lv.slotCap ? n : lv.slotCap;
        const int sub = lane & 7;
        for (int i0 = warp * 4; i0 < nOut; i0 += (OCT_T / 32) * 4) {
            const int i = i0 + (lane >> 3);
            unsigned long long k = 0;
            if (i < nOut) {
                const uint32_t node = cur[i];
                NODE_DECODE(node);
                for (int y = y0_ + sub; y < y1_; y += 8) {
                    const unsigned long long v = bestS[pb_ + y];
                    k = v > k ? v : k;
                }
            }
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                const unsigned long long v = __shfl_xor_sync(0xffffffffu, k, o);
                k = v > k ? v : k;
            }
            if (sub == 0 && i < nOut) {
                const int x = (int)((k >> 14) & 0x3fff), y = (int)(k & 0x3fff), sc = (int)(k >> 56);
                out[i] = make_int2(x | (y << 16), sc);
            }
        }
        if (tid == 0) lvlCount[frame * L.nlevels + level] = nOut;
    }
#undef NODE_DECODE
#undef NODE_SIZE
}

size_t octree_smem_bytes(int maxRows, int maxNodes)
{
    return octree_lists_bytes(maxRows, maxNodes) + (size_t)maxRows * 8;
}

cudaError_t launch_octree(const OrbxLayout &L, const uint32_t *cnt, const unsigned long long *best, int2 *slots,
                          int *lvlCount, int maxRows, int maxNodes, int batch, cudaStream_t st)
{
    const size_t smem = octree_smem_bytes(maxRows, maxNodes);
    // opt in per call (the attribute is per device and this is a cheap host-side set).  The 48 KB default limit covers static +
    // dynamic shared memory together, so the opt-in starts well below it: a dynamic size just under 48 KB plus the kernel's
    // few hundred static bytes is an invalid launch otherwise.
    {
        cudaError_t e = orbx_raise_dyn_smem((const void *)k_octree, smem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid(L.nlevels, batch);
    k_octree<<<grid, OCT_T, smem, st>>>(L, cnt, best, slots, lvlCount, maxRows, maxNodes);
    return cudaGetLastError();
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

#define DS_WARPS 4          // warps per CTA
#define DS_PER_WARP 8       // keypoint slots handled by one warp, one after the other
#define DS_PB 64            // blurred patch: TMA box 64 x 37 bytes (16-byte aligned start, 37 columns from offset 0..15)
#define DS_PA 48            // unblurred patch: TMA box 48 x 31 bytes (32 columns from offset 0..15)
#define DS_BUFB (37 * DS_PB + 64)          // 2432: keeps the second box 128-byte aligned
#define DS_BUF (DS_BUFB + 31 * DS_PA + 48) // 3968 bytes per stage
#define DS_TX (37 * DS_PB + 31 * DS_PA)    // bytes the two box loads of one keypoint deliver

__device__ __forceinline__ int dp4a_u8_s8(uint32_t pix, uint32_t wgt, int acc)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(pix), "r"(wgt), "r"(acc));
    return d;
}

struct DescSlot { int valid, cx, cy, level, out, score; };

__global__ void __launch_bounds__(DS_WARPS * 32)
k_describe(const __grid_constant__ OrbxTensorMaps tmA, const __grid_constant__ OrbxTensorMaps tmB, int f0, const __grid_constant__ OrbxLayout L,
           const int2 *__restrict__ slots, const int *__restrict__ lvlCount, DescUmax um,
           orbx_keypoint_pod *__restrict__ kps, uint8_t *__restrict__ desc, int *__restrict__ counts)
{
    // per warp two stages of {blurred 37x37 patch (rBRIEF samples), unblurred 31x31 patch (IC_Angle)}: the
    // patches of the warp's next keypoint arrive by two TMA box loads (one elected lane, completion on the
    // stage's mbarrier) while the current one is being processed
    __shared__ __align__(128) uint8_t patch[DS_WARPS][2][DS_BUF];
    __shared__ __align__(8) uint64_t bars[DS_WARPS][2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.y;
    // per-level counts of this frame -> exclusive prefix (lane l holds level l)
    const int *lc = lvlCount + frame * L.nlevels;
    const int myCnt = lane < L.nlevels ? lc[lane] : 0;
    int incl = myCnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (blockIdx.x == 0 && tid == 0) counts[frame] = total;
    const int myBase = lane < L.nlevels ? L.lv[lane].slotBase : 0x7fffffff;
    if (lane < 2) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[warp][lane])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();

    // lane j < DS_PER_WARP fetches the warp's j-th slot record now: one global load per warp instead of one, with its latency,
    // in front of every key point's TMA requests
    int2 mySlot = make_int2(0, 0);
    {
        const int slot = (blockIdx.x * DS_PER_WARP + lane) * DS_WARPS + warp;
        if (lane < DS_PER_WARP && slot < L.slotsPerFrame) mySlot = __ldg(&slots[(size_t)frame * L.slotsPerFrame + slot]);
    }
    // resolve slot j of this warp and start streaming its two patches into stage `buf`
    auto issue = [&](int j, int buf) -> DescSlot {
        DescSlot s; s.valid = 0; s.cx = s.cy = s.level = s.out = s.score = 0;
        const int slot = (blockIdx.x * DS_PER_WARP + j) * DS_WARPS + warp;     // warp-uniform
        if (j >= DS_PER_WARP || slot >= L.slotsPerFrame) return s;
        const unsigned ge = __ballot_sync(0xffffffffu, slot >= myBase);        // level = last l with slotBase[l] <= slot
        const int level = 31 - __clz((int)ge);
        const OrbxLevel &lv = L.lv[level];
        const int i = slot - lv.slotBase;
        const int cntL = __shfl_sync(0xffffffffu, myCnt, level);
        const int before = __shfl_sync(0xffffffffu, incl - myCnt, level);
        if (i >= cntL) return s;
        const int2 sl = make_int2(__shfl_sync(0xffffffffu, mySlot.x, j), __shfl_sync(0xffffffffu, mySlot.y, j));
        s.valid = 1; s.level = level; s.out = before + i; s.score = sl.y;
        s.cx = (sl.x & 0xffff) + ORBX_MINB; s.cy = (sl.x >> 16) + ORBX_MINB;   // :984-985
        if (lane == 0) {
            // box starts are 16-byte aligned: the patch column cx-18 (cx-15) sits at byte (cx-18) & 15 ((cx-15) & 15) of its row
            uint8_t *sB = patch[warp][buf], *sA = sB + DS_BUFB;
            const uint32_t bb = smem_u32(&bars[warp][buf]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bb), "r"((uint32_t)DS_TX) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_u32(sB)), "l"(&tmB.m[level]), "r"((s.cx - 18) & ~15), "r"(s.cy - 18), "r"(f0 + frame), "r"(bb) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_u32(sA)), "l"(&tmA.m[level]), "r"((s.cx - 15) & ~15), "r"(s.cy - 15), "r"(f0 + frame), "r"(bb) : "memory");
        }
        return s;
    };

    // the first key point's patches are requested before the lane's constants are set up: their latency overlaps the set-up
    DescSlot cur = issue(0, 0);
    // this lane's 16 pattern points (descriptor byte `lane`), kept in registers across its slots
    float px[16], py[16];
    {
        const int4 a = __ldg((const int4 *)(d_pattern + lane * 32)), b = __ldg((const int4 *)(d_pattern + lane * 32 + 16));
        const int w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int k = 0; k < 8; k++) {
            px[2 * k] = (float)(int8_t)(w[k] & 0xff);         py[2 * k] = (float)(int8_t)((w[k] >> 8) & 0xff);
            px[2 * k + 1] = (float)(int8_t)((w[k] >> 16) & 0xff); py[2 * k + 1] = (float)(int8_t)((w[k] >> 24) & 0xff);
        }
    }
    // IC_Angle weights of this lane's patch row v = lane-15: byte j of word k is column u = 4k+j-15;
    // wu = u inside the disc (|u| <= umax[|v|], orbextractor.cpp:150-152), w1 = 1 inside, 0 outside
    uint32_t wu[8], w1[8];
    {
        const int v = lane - 15, av = v < 0 ? -v : v;
        int d = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) if (k == av) d = um.u[k];
        if (lane == 31) d = -1;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            uint32_t a = 0, b = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int u = 4 * k + j - 15;
                const bool in = (u <= d) && (-u <= d);
                a |= in ? ((uint32_t)(u & 0xff) << (8 * j)) : 0u;
                b |= in ? (1u << (8 * j)) : 0u;
            }
            wu[k] = a; w1[k] = b;
        }
    }
    unsigned phase = 0;                                       // bit b = parity the next completion of stage b's barrier has
    for (int j = 0; j < DS_PER_WARP; j++) {
        const int buf = j & 1;
        const DescSlot nxt = issue(j + 1, buf ^ 1);
        if (cur.valid) {
            mbar_wait(&bars[warp][buf], (phase >> buf) & 1u);
            phase ^= 1u << buf;
            const OrbxLevel &lv = L.lv[cur.level];
            const uint8_t *pB = patch[warp][buf], *pA = pB + DS_BUFB;
            const int shB = (cur.cx - 18) & 15, shA = (cur.cx - 15) & 15;
            // ---- IC_Angle: lane = patch row v; 32 bytes of the row (cols cx-15 .. cx+16) against the weights
            int m10 = 0, rowsum = 0;
            if (lane < 31) {
                const uint32_t *rw = (const uint32_t *)(pA + lane * DS_PA + (shA & ~3));
                uint32_t W[9];
#pragma unroll
                for (int k = 0; k < 9; k++) W[k] = rw[k];
                const int sh = (shA & 3) * 8;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const uint32_t B = __funnelshift_r(W[k], W[k + 1], sh);
                    m10 = dp4a_u8_s8(B, wu[k], m10);
                    rowsum = dp4a_u8_s8(B, w1[k], rowsum);
                }
            }
            // warp sums in one instruction each (REDUX; integer addition, so the order of the terms is immaterial)
            m10 = __reduce_add_sync(0xffffffffu, m10);
            const int m01 = __reduce_add_sync(0xffffffffu, (lane - 15) * rowsum);
            const float angle = fast_atan2_deg((float)m01, (float)m10);

            // ---- rBRIEF: lane = descriptor byte, 16 rotated samples each, gathered from shared memory.
            // cvRound = round-half-even: adding 1.5 * 2^23 leaves exactly that integer in the low mantissa bits.
            const float factorPI = (float)(3.1415926535897932384626433832795 / 180.0);
            float sa, ca;
            glibc_sincosf(__fmul_rn(angle, factorPI), &sa, &ca);
            const float a = ca, b = sa;
            const float MAGIC = 12582912.0f;                       // 0x4B400000
            // shared-window address of the patch centre with both magic offsets taken off once (mod 2^32): a sample's address is
            // row_bits * DS_PB + col_bits + base32, two integer instructions
            const uint32_t base32 = smem_u32(pB + 18 * DS_PB + shB + 18) - 0x4B400000u * (unsigned)(DS_PB + 1);
            int val = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                int t[2];
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const float x = px[2 * k + e], y = py[2 * k + e];
                    const int row = __float_as_int(__fadd_rn(__fadd_rn(__fmul_rn(x, b), __fmul_rn(y, a)), MAGIC));
                    const int col = __float_as_int(__fadd_rn(__fsub_rn(__fmul_rn(x, a), __fmul_rn(y, b)), MAGIC));
                    unsigned v;
                    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"((unsigned)row * (unsigned)DS_PB + (unsigned)col + base32));
                    t[e] = (int)v;
                }
                val |= (t[0] < t[1]) << k;
            }
            desc[((size_t)frame * L.kpStride + cur.out) * 32 + lane] = (uint8_t)val;
            if (lane == 0) {
                orbx_keypoint_pod kp;
                float fx = (float)cur.cx, fy = (float)cur.cy;
                if (cur.level != 0) { fx = __fmul_rn(fx, lv.sf); fy = __fmul_rn(fy, lv.sf); }
                kp.x = fx; kp.y = fy; kp.size = (float)lv.kpSize; kp.angle = angle; kp.response = (float)cur.score;
                kp.octave = cur.level; kp.class_id = -1;
                kps[(size_t)frame * L.kpStride + cur.out] = kp;
            }
        }
        __syncwarp();     // stage `buf` is refilled two iterations from now
        cur = nxt;
    }
}

void launch_describe(const OrbxTensorMaps &mapsA, const OrbxTensorMaps &mapsB, int f0, const OrbxLayout &L, const int2 *slots,
                     const int *lvlCount, const int umax[16], orbx_keypoint_pod *kps, uint8_t *desc, int *counts,
                     int batch, cudaStream_t st)
{
    DescUmax um;
    for (int k = 0; k < 16; k++) um.u[k] = umax[k];
    const int perBlock = DS_WARPS * DS_PER_WARP;
    dim3 grid((L.slotsPerFrame + perBlock - 1) / perBlock, batch);
    k_describe<<<grid, DS_WARPS * 32, 0, st>>>(mapsA, mapsB, f0, L, slots, lvlCount, um, kps, desc, counts);
}

// ------------------------------------------------------------------------------------------
// OrbFrame::ComputeStereoMatches (orbframe.cpp:511-705, "next" row N1 of SURVEY 8f): the immediate
// consumer of both extractors' keypoints, descriptors and pyramids -- run here on the device-resident
// results so neither pyramid has to leave HBM.
//   k_stereo_match : one warp per left keypoint.  Candidates are the right keypoints whose row band
//                    [floor(y-r), ceil(y+r)], r = 2*scale[octave], contains the left row (:527-541, the
//                    reference's vRowIndices table, evaluated on the fly), within one octave and inside
//                    the disparity range; best Hamming distance below TH_HIGH, lowest index on ties
//                    (:565-597); then the 11x11 SAD over 11 offsets at the keypoint's level and the
//                    parabola fit (:600-690), float32 without FMA.
//   k_stereo_filter: one CTA per pair: sort (SAD, index), median, reject SAD >= 1.5*1.4*median (:693-705).

// ------------------------------------------------------------------------------------------
// OrbFrame::FilterKeyPoints (orbframe.cpp:403-445): key points strictly inside the bounding box are dropped, the others keep
// their order; key points and descriptors of one frame are compacted in place.  One CTA per frame walks the frame in chunks
// of 1024: a chunk is read into registers, a barrier, then written -- every write lands at or left of the chunk's own
// positions, which have all been read by then.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_filter_keypoints(orbx_keypoint_pod *__restrict__ kps, uint8_t *__restrict__ desc, int *__restrict__ counts, int kpStride,
                   int frame0, float bx0, float bx1, float by0, float by1)
{
    __shared__ int wsum[32];
    __shared__ int carry;
    const int frame = frame0 + blockIdx.x;
    orbx_keypoint_pod *k = kps + (size_t)frame * kpStride;
    uint4 *d = (uint4 *)(desc + (size_t)frame * kpStride * 32);
    const int n = counts[frame];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 1024) {
        const int i = i0 + tid;
        orbx_keypoint_pod kp = {};
        uint4 da = make_uint4(0, 0, 0, 0), db = da;
        bool keep = false;
        if (i < n) {
            kp = k[i]; da = d[2 * i]; db = d[2 * i + 1];
            const bool inBounds = kp.x > bx0 && kp.x < bx1 && kp.y > by0 && kp.y < by1;   // :411-412
            keep = !inBounds;                                                             // :413
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();                       // every entry of the chunk is in registers
        int base = carry;
        for (int w = 0; w < warp; w++) base += wsum[w];
        if (keep) {
            const int o = base + __popc(bal & ((1u << lane) - 1u));
            k[o] = kp; d[2 * o] = da; d[2 * o + 1] = db;
        }
        __syncthreads();
        if (tid == 0) { int t = carry; for (int w = 0; w < 32; w++) t += wsum[w]; carry = t; }
        __syncthreads();
    }
    if (tid == 0) counts[frame] = carry;
}

cudaError_t launch_filter_keypoints(orbx_keypoint_pod *kps, uint8_t *desc, int *counts, int kpStride, int frame0, int nFrames,
                                    const float box[4], cudaStream_t st)
{
    k_filter_keypoints<<<nFrames, 1024, 0, st>>>(kps, desc, counts, kpStride, frame0, box[0], box[1], box[2], box[3]);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int hamming256_dev(const uint4 &qa, const uint4 &qb, const uint4 &a, const uint4 &b)
{
    return __popc(qa.x ^ a.x) + __popc(qa.y ^ a.y) + __popc(qa.z ^ a.z) + __popc(qa.w ^ a.w) +
           __popc(qb.x ^ b.x) + __popc(qb.y ^ b.y) + __popc(qb.z ^ b.z) + __popc(qb.w ^ b.w);
}

__global__ void __launch_bounds__(128)
k_stereo_match(const __grid_constant__ OrbxLayout L, const uint8_t *__restrict__ pyrL, const uint8_t *__restrict__ pyrR,
               const orbx_keypoint_pod *__restrict__ kl, const uint4 *__restrict__ dl, const int *__restrict__ nlPtr,
               const orbx_keypoint_pod *__restrict__ kr, const uint4 *__restrict__ dr, const int *__restrict__ nrPtr,
               int frameStep, float mbf, float maxD, float *__restrict__ uRight, float *__restrict__ depth, int *__restrict__ sad)
{
    __shared__ float part[4][121];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int iL = blockIdx.x * 4 + warp;
    {   // pair p = blockIdx.y: both sides advance by frameStep frames, the outputs by one record row
        const size_t fo = (size_t)blockIdx.y * frameStep;
        pyrL += fo * L.slab; pyrR += fo * L.slab;
        kl += fo * L.kpStride; kr += fo * L.kpStride; dl += fo * L.kpStride * 2; dr += fo * L.kpStride * 2;
        nlPtr += fo; nrPtr += fo;
        uRight += (size_t)blockIdx.y * L.kpStride; depth += (size_t)blockIdx.y * L.kpStride; sad += (size_t)blockIdx.y * L.kpStride;
    }
    const int nl = *nlPtr, nr = *nrPtr;
    if (iL >= nl) return;
    const orbx_keypoint_pod kpL = kl[iL];
    if (lane == 0) { uRight[iL] = -1.0f; depth[iL] = -1.0f; sad[iL] = -1; }
    const int levelL = kpL.octave;
    const float vL = kpL.y, uL = kpL.x;
    const int row = (int)vL;
    const float minU = __fsub_rn(uL, maxD), maxU = uL;           // minD = 0
    if (maxU < 0) return;
    const uint4 qa = __ldg(&dl[2 * iL]), qb = __ldg(&dl[2 * iL + 1]);
    const int TH_HIGH = 100, thOrbDist = 75;
    unsigned best = (unsigned)TH_HIGH << 16;                     // dist < TH_HIGH only; lowest iR wins ties
    for (int iR = lane; iR < nr; iR += 32) {
        const orbx_keypoint_pod k = kr[iR];
        const float r = __fmul_rn(2.0f, L.lv[k.octave].sf);
        const int maxr = (int)ceilf(__fadd_rn(k.y, r)), minr = (int)floorf(__fsub_rn(k.y, r));
        if (row < minr || row > maxr) continue;
        if (k.octave < levelL - 1 || k.octave > levelL + 1) continue;
        if (!(k.x >= minU && k.x <= maxU)) continue;
        const int d = hamming256_dev(qa, qb, __ldg(&dr[2 * iR]), __ldg(&dr[2 * iR + 1]));
        const unsigned key = (unsigned)d << 16 | (unsigned)iR;
        best = min(best, key);
    }
    best = __reduce_min_sync(0xffffffffu, best);
    const int bestDist = (int)(best >> 16), bestIdxR = (int)(best & 0xffff);
    if (bestDist >= thOrbDist) return;

    // ---- sub-pixel match by correlation, at the left keypoint's pyramid level
    const OrbxLevel &lv = L.lv[levelL];
    const float uR0 = kr[bestIdxR].x;
    const float scaleFactor = lv.invSf;
    const float scaleduL = roundf(__fmul_rn(kpL.x, scaleFactor)), scaledvL = roundf(__fmul_rn(kpL.y, scaleFactor));
    const float scaleduR0 = roundf(__fmul_rn(uR0, scaleFactor));
    const int w = 5, LL = 5;
    const float iniu = scaleduR0 + LL - w, endu = scaleduR0 + LL + w + 1;
    if (iniu < 0 || endu >= lv.w) return;
    const int y0 = (int)(scaledvL - w), xl0 = (int)(scaleduL - w), xr0 = (int)(scaleduR0 - w);
    // the reference would throw inside cv::Mat::colRange / rowRange here; no match is reported instead
    if (y0 < 0 || y0 + 11 > lv.h || xl0 < 0 || xl0 + 11 > lv.w || xr0 - LL < 0 || xr0 + LL + 11 > lv.w) return;
    const uint8_t *IL = pyrL + lv.off + (size_t)y0 * lv.pitch + xl0;
    const uint8_t *IR = pyrR + lv.off + (size_t)y0 * lv.pitch + xr0;
    const int cL = IL[w * lv.pitch + w];
    // item k = (offset index, window row): 121 items over 32 lanes; the centre of the right window moves with incR
    for (int k = lane; k < 121; k += 32) {
        const int inc = k / 11, ry = k - inc * 11;
        const uint8_t *a = IL + ry * lv.pitch, *b = IR + ry * lv.pitch + (inc - LL);
        const int cR = IR[w * lv.pitch + w + (inc - LL)];
        int s = 0;
#pragma unroll
        for (int x = 0; x < 11; x++) { const int v = ((int)a[x] - cL) - ((int)b[x] - cR); s += v < 0 ? -v : v; }
        part[warp][k] = (float)s;
    }
    __syncwarp();
    float dist = 0.f;
    if (lane < 11) {
#pragma unroll
        for (int ry = 0; ry < 11; ry++) dist += part[warp][lane * 11 + ry];   // integers < 2^24: exact in float
    }
    float vd[11];
#pragma unroll
    for (int i = 0; i < 11; i++) vd[i] = __shfl_sync(0xffffffffu, dist, i);
    if (lane != 0) return;
    int bestSad = 0x7fffffff, bestinc = 0;
#pragma unroll
    for (int i = 0; i < 11; i++)
        if (vd[i] < (float)bestSad) { bestSad = (int)vd[i]; bestinc = i - LL; }
    if (bestinc == -LL || bestinc == LL) return;
    float d1 = 0, d2 = 0, d3 = 0;
#pragma unroll
    for (int i = 1; i < 10; i++) if (i == bestinc + LL) { d1 = vd[i - 1]; d2 = vd[i]; d3 = vd[i + 1]; }
    const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));
    if (deltaR < -1 || deltaR > 1) return;
    float bestuR = __fmul_rn(lv.sf, __fadd_rn(__fadd_rn(scaleduR0, (float)bestinc), deltaR));
    float disparity = __fsub_rn(uL, bestuR);
    if (disparity >= 0 && disparity < maxD) {
        if (disparity <= 0) { disparity = 0.01f; bestuR = (float)((double)uL - 0.01); }
        depth[iL] = __fdiv_rn(mbf, disparity);
        uRight[iL] = bestuR;
        sad[iL] = bestSad;
    }
}

__global__ void __launch_bounds__(256)
k_stereo_filter(const int *__restrict__ nlPtr, int frameStep, int kpStride, const int *__restrict__ sad, float *__restrict__ uRight,
                float *__restrict__ depth, int *__restrict__ nMatches, int pow2)
{
    extern __shared__ unsigned skeys[];
    __shared__ int cnt;
    nlPtr += (size_t)blockIdx.x * frameStep; nMatches += blockIdx.x;
    sad += (size_t)blockIdx.x * kpStride; uRight += (size_t)blockIdx.x * kpStride; depth += (size_t)blockIdx.x * kpStride;
    const int nl = *nlPtr, tid = threadIdx.x;
    if (tid == 0) cnt = 0;
    __syncthreads();
    for (int i = tid; i < pow2; i += 256) {
        unsigned key = 0xffffffffu;
        if (i < nl && sad[i] >= 0) { key = (unsigned)sad[i] << 16 | (unsigned)i; atomicAdd(&cnt, 1); }
        skeys[i] = key;
    }
    __syncthreads();
    const int n = cnt;
    if (tid == 0) *nMatches = n;
    if (n == 0) return;
    for (int k = 2; k <= pow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < pow2; t += 256) {
                const int x = t ^ j;
                if (x > t) {
                    const unsigned a = skeys[t], b = skeys[x];
                    const bool asc = (t & k) == 0;
                    if ((a > b) == asc) { skeys[t] = b; skeys[x] = a; }
                }
            }
            __syncthreads();
        }
    const float median = (float)(skeys[n / 2] >> 16);
    const float thDist = __fmul_rn(__fmul_rn(1.5f, 1.4f), median);
    for (int i = tid; i < n; i += 256) {
        const unsigned key = skeys[i];
        if (!((float)(key >> 16) < thDist)) { uRight[key & 0xffff] = -1.0f; depth[key & 0xffff] = -1.0f; }
    }
}

// nPairs pairs: pair p reads the frames p * frameStep after the given left / right pointers and writes row p of
// uRight / depth / sad (kpStride entries each) and nMatches[p]
cudaError_t launch_stereo(const OrbxLayout &L, const uint8_t *pyrL, const uint8_t *pyrR, const orbx_keypoint_pod *kl,
                          const uint8_t *dl, const int *nl, const orbx_keypoint_pod *kr, const uint8_t *dr, const int *nr,
                          int nPairs, int frameStep, float mbf, float maxD, float *uRight, float *depth, int *sad, int *nMatches, cudaStream_t st)
{
    const int cap = L.kpStride;
    if (cap > 65535) return cudaErrorInvalidValue;
    int pow2 = 2; while (pow2 < cap) pow2 <<= 1;
    k_stereo_match<<<dim3((cap + 3) / 4, nPairs), 128, 0, st>>>(L, pyrL, pyrR, kl, (const uint4 *)dl, nl, kr, (const uint4 *)dr, nr, frameStep,
                                                                 mbf, maxD, uRight, depth, sad);
    const size_t smem = (size_t)pow2 * sizeof(unsigned);
    {   // more than 8192 key points per frame need the opt-in
        cudaError_t e = orbx_raise_dyn_smem((const void *)k_stereo_filter, smem);
        if (e != cudaSuccess) return e;
    }
    k_stereo_filter<<<nPairs, 256, smem, st>>>(nl, frameStep, cap, sad, uRight, depth, nMatches, pow2);
    return cudaGetLastError();
}
