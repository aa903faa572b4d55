// standalone probe: u8 TMA box loads as k_blur issues them; argv[1] selects the variant
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#define SMEM_BYTES (144 * 134)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ void body(const CUtensorMap *map, int rank, int x, int y, int z, uint32_t bytes, uint32_t *out)
{
    __shared__ __align__(128) uint8_t tileS[SMEM_BYTES];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t b = smem_u32(&bar), d = smem_u32(tileS);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        if (rank == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(d), "l"(map), "r"(x), "r"(y), "r"(z), "r"(b) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(d), "l"(map), "r"(x), "r"(y), "r"(b) : "memory");
    }
    const uint32_t b = smem_u32(&bar);
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra WAIT_%=;\n\t}" ::"r"(b), "r"(0) : "memory");
    uint32_t s = 0;
    for (int i = threadIdx.x; i < (int)bytes; i += blockDim.x) s += tileS[i];
    atomicAdd(out, s);
}
__global__ void k_global(const CUtensorMap *map, int rank, int x, int y, int z, uint32_t bytes, uint32_t *out) { body(map, rank, x, y, z, bytes, out); }
__global__ void k_param(const __grid_constant__ CUtensorMap map, int rank, int x, int y, int z, uint32_t bytes, uint32_t *out) { body(&map, rank, x, y, z, bytes, out); }
int main(int argc, char **argv)
{
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    int w = 1241, h = 376, pitch = 1280, frames = 4; size_t slab = (size_t)pitch * h;
    uint8_t *d; cudaMalloc(&d, slab * frames + 512); cudaMemset(d, 1, slab * frames + 512);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeTiledFn encode = (EncodeTiledFn)fn;
    int rank = (variant == 2) ? 2 : 3;
    cuuint32_t bw = (variant == 3 || variant == 5) ? 128 : 144, bh = (variant == 3) ? 64 : 134;
    if (variant == 6) { bw = 144; bh = 64; }
    if (variant == 7) { bw = 256; bh = 64; }
    CUtensorMap hm;
    cuuint64_t gdim[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)frames};
    cuuint64_t gstr[2] = {(cuuint64_t)pitch, (cuuint64_t)slab};
    cuuint32_t box[3] = {bw, bh, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&hm, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d + 256, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUtensorMap *dm; cudaMalloc(&dm, sizeof(hm)); cudaMemcpy(dm, &hm, sizeof(hm), cudaMemcpyHostToDevice);
    uint32_t *out; cudaMalloc(&out, 4); cudaMemset(out, 0, 4);
    int x = (variant == 4) ? 0 : -4, y = (variant == 4) ? 0 : -3;
    if (variant == 1) k_param<<<1, 128>>>(hm, rank, x, y, 1, bw * bh, out);
    else k_global<<<1, 128>>>(dm, rank, x, y, 1, bw * bh, out);
    cudaError_t e = cudaDeviceSynchronize();
    uint32_t ho = 0; cudaMemcpy(&ho, out, 4, cudaMemcpyDeviceToHost);
    printf("variant %d rank %d box %ux%u encode=%d: %s sum=%u\n", variant, rank, bw, bh, (int)r, cudaGetErrorString(e), ho);
    return 0;
}
