// orbv_vocab.cu -- "next" row N2 of SURVEY 8(f): the bag-of-words descent, pure Hamming work on descriptors:
//   k_voc_transform   OrbVocabulary::transform5                    orbvocabulary.cpp:203-242
// Distances are OrbDescriptor::distance (orbdescriptor.cpp:75-95) == ORBmatcher::DescriptorDistance
// (orbmatcher.cpp:1662-1677): popcount of the XOR of two 256-bit rows.
#include "../../include/orbx.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

// ------------------------------------------------------------------------------------------
// Vocabulary descent: 8 lanes per feature, lane s holds word s of the feature and of every child it is
// compared with (one coalesced 32-byte read per child and group); the distance is a 3-step butterfly
// inside the group.  The first child with the least distance is followed (strict '<', :224-232).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_voc_transform(const int *__restrict__ childOff, const int *__restrict__ childIds, const uint32_t *__restrict__ nodeDesc,
                const int *__restrict__ wordId, int nodeLevel, const uint32_t *__restrict__ feat, size_t featStrideWords, int n,
                int2 *__restrict__ out)
{
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, sub = threadIdx.x & 7;
    if (g >= n) return;
    const unsigned mask = 0xffu << (threadIdx.x & 24);
    const uint32_t q = __ldg(&feat[(size_t)g * featStrideWords + sub]);
    int node = 0, level = 0, nid = 0;
    int off = __ldg(&childOff[0]), nc = __ldg(&childOff[1]) - off;
    while (nc > 0) {
        ++level;
        int best = 0x7fffffff, bestId = 0;
        for (int c = 0; c < nc; c++) {
            const int id = __ldg(&childIds[off + c]);
            int d = __popc(q ^ __ldg(&nodeDesc[(size_t)id * 8 + sub]));
            d += __shfl_xor_sync(mask, d, 1);
            d += __shfl_xor_sync(mask, d, 2);
            d += __shfl_xor_sync(mask, d, 4);
            if (d < best) { best = d; bestId = id; }
        }
        node = bestId;
        if (level == nodeLevel) nid = node;
        off = __ldg(&childOff[node]); nc = __ldg(&childOff[node + 1]) - off;
    }
    if (sub == 0) out[g] = make_int2(__ldg(&wordId[node]), nid);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct orbv_vocab {
    int device = 0, nNodes = 0, L = 0;
    cudaStream_t stream = nullptr;
    int *dChildOff = nullptr, *dChildIds = nullptr, *dWordId = nullptr;
    uint32_t *dNodeDesc = nullptr;
    std::vector<int32_t> wordId;
    std::vector<double> weight;
    uint8_t *dFeat = nullptr; int2 *dOut = nullptr; int2 *hOut = nullptr; int cap = 0;
    std::string err;
};

namespace {
int vfail(orbv_vocab *v, int code, const std::string &msg) { if (v) v->err = msg; return code; }
#define VCK(call)                                                                                     \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) return vfail(v, ORBX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)
} // namespace

extern "C" {

int orbv_create(int device, int n_nodes, const int32_t *child_off, const int32_t *child_ids, const uint8_t *node_desc,
                const int32_t *word_id, const double *weight, int L, orbv_vocab **out)
{
    if (!out) return ORBX_ERR_ARG;
    *out = nullptr;
    if (n_nodes < 1 || !child_off || !child_ids || !node_desc || !word_id || !weight || L < 1) return ORBX_ERR_ARG;
    orbv_vocab *v = new (std::nothrow) orbv_vocab();
    if (!v) return ORBX_ERR_NOMEM;
    *out = v;
    v->device = device; v->nNodes = n_nodes; v->L = L;
    // the tree must be walkable: the root has children (transform5 reads children[0] unconditionally, :220-221),
    // child ids are nodes, offsets are monotone
    if (child_off[0] != 0 || child_off[1] <= 0) return vfail(v, ORBX_ERR_ARG, "the root needs children");
    for (int i = 0; i < n_nodes; i++) if (child_off[i + 1] < child_off[i]) return vfail(v, ORBX_ERR_ARG, "child offsets must be non-decreasing");
    const int nnz = child_off[n_nodes];
    for (int k = 0; k < nnz; k++) if (child_ids[k] <= 0 || child_ids[k] >= n_nodes) return vfail(v, ORBX_ERR_ARG, "child id out of range");
    int devCount = 0;
    cudaError_t e = cudaGetDeviceCount(&devCount);
    if (e != cudaSuccess || device < 0 || device >= devCount) { cudaGetLastError(); return vfail(v, ORBX_ERR_CUDA, "no such CUDA device"); }
    VCK(cudaSetDevice(device));
    cudaDeviceProp prop;
    VCK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return vfail(v, ORBX_ERR_CUDA, "liborbx is built for sm_100a only (no other code path exists)");
    VCK(cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking));
    VCK(cudaMalloc((void **)&v->dChildOff, sizeof(int) * (size_t)(n_nodes + 1)));
    VCK(cudaMalloc((void **)&v->dChildIds, sizeof(int) * (size_t)std::max(nnz, 1)));
    VCK(cudaMalloc((void **)&v->dWordId, sizeof(int) * (size_t)n_nodes));
    VCK(cudaMalloc((void **)&v->dNodeDesc, (size_t)n_nodes * 32));
    VCK(cudaMemcpy(v->dChildOff, child_off, sizeof(int) * (size_t)(n_nodes + 1), cudaMemcpyHostToDevice));
    if (nnz > 0) VCK(cudaMemcpy(v->dChildIds, child_ids, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice));
    VCK(cudaMemcpy(v->dWordId, word_id, sizeof(int) * (size_t)n_nodes, cudaMemcpyHostToDevice));
    VCK(cudaMemcpy(v->dNodeDesc, node_desc, (size_t)n_nodes * 32, cudaMemcpyHostToDevice));
    v->wordId.assign(word_id, word_id + n_nodes);
    int nWords = 0;
    for (int i = 0; i < n_nodes; i++) nWords = std::max(nWords, word_id[i] + 1);
    v->weight.assign((size_t)std::max(nWords, 1), 0.0);
    for (int i = 0; i < n_nodes; i++) if (word_id[i] >= 0) v->weight[word_id[i]] = weight[i];
    return ORBX_OK;
}

void orbv_destroy(orbv_vocab *v)
{
    if (!v) return;
    if (v->stream) { cudaSetDevice(v->device); cudaStreamSynchronize(v->stream); cudaStreamDestroy(v->stream); }
    cudaFree(v->dChildOff); cudaFree(v->dChildIds); cudaFree(v->dWordId); cudaFree(v->dNodeDesc);
    cudaFree(v->dFeat); cudaFree(v->dOut);
    if (v->hOut) cudaFreeHost(v->hOut);
    delete v;
}

const char *orbv_last_error(const orbv_vocab *v) { return v ? v->err.c_str() : "null handle"; }

int orbv_transform_device(orbv_vocab *v, const uint8_t *d_desc, size_t desc_stride, int n, int levels_up, int32_t *d_word_node, void *stream)
{
    if (!v) return ORBX_ERR_ARG;
    if (!d_desc || !d_word_node || n < 1 || desc_stride < 32 || (desc_stride & 3)) return vfail(v, ORBX_ERR_ARG, "bad argument");
    VCK(cudaSetDevice(v->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : v->stream;
    const int groupsPerCta = 128 / 8;
    k_voc_transform<<<(n + groupsPerCta - 1) / groupsPerCta, 128, 0, st>>>(v->dChildOff, v->dChildIds, v->dNodeDesc, v->dWordId,
                                                                            v->L - levels_up, (const uint32_t *)d_desc, desc_stride / 4, n,
                                                                            (int2 *)d_word_node);
    VCK(cudaGetLastError());
    return ORBX_OK;
}

int orbv_transform(orbv_vocab *v, const uint8_t *desc, int n, int levels_up, int32_t *word_id, double *weight, int32_t *node_id)
{
    if (!v) return ORBX_ERR_ARG;
    if (!desc || !word_id || n < 1) return vfail(v, ORBX_ERR_ARG, "bad argument");
    VCK(cudaSetDevice(v->device));
    if (n > v->cap) {
        cudaFree(v->dFeat); cudaFree(v->dOut); if (v->hOut) cudaFreeHost(v->hOut);
        v->dFeat = nullptr; v->dOut = nullptr; v->hOut = nullptr; v->cap = 0;
        VCK(cudaMalloc((void **)&v->dFeat, (size_t)n * 32));
        VCK(cudaMalloc((void **)&v->dOut, (size_t)n * sizeof(int2)));
        VCK(cudaMallocHost((void **)&v->hOut, (size_t)n * sizeof(int2)));
        v->cap = n;
    }
    VCK(cudaMemcpyAsync(v->dFeat, desc, (size_t)n * 32, cudaMemcpyHostToDevice, v->stream));
    int rc = orbv_transform_device(v, v->dFeat, 32, n, levels_up, (int32_t *)v->dOut, v->stream);
    if (rc != ORBX_OK) return rc;
    VCK(cudaMemcpyAsync(v->hOut, v->dOut, (size_t)n * sizeof(int2), cudaMemcpyDeviceToHost, v->stream));
    VCK(cudaStreamSynchronize(v->stream));
    for (int i = 0; i < n; i++) {
        word_id[i] = v->hOut[i].x;
        if (node_id) node_id[i] = v->hOut[i].y;
        if (weight) weight[i] = v->hOut[i].x >= 0 && (size_t)v->hOut[i].x < v->weight.size() ? v->weight[v->hOut[i].x] : 0.0;
    }
    return ORBX_OK;
}

} // extern "C"
