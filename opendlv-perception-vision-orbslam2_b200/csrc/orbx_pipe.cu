// orbx_pipe.cu -- several device-resident extractions in flight on one GPU.
//
// One orbx_extract_batch_device call fills and drains the GPU: the first stages of a batch (level-0 copy, the chain of seven
// dependent resize launches) and its last ones (octree: a few latency-bound CTAs; the tail of the describe grid) leave most SMs
// idle, and the next call on the same handle cannot start before the previous one has released the handle's arenas.  A pipe owns
// `depth` ordinary extractor handles -- each with its own pyramid / blur / result arenas and its own stream -- and hands
// consecutive submissions to them in turn, so the head of submission k+1 runs under the tail of submission k (measured on B200:
// 0.415 -> 0.372 -> 0.365 ms per 64-frame KITTI batch with 1 / 2 / 3 in flight).  This is the device-side twin of
// orbx_extract_batch_async / orbx_wait; the reference has the same shape (extract in threads, consume later,
// orbframe.cpp:73-78).  Built on the public C ABI only.
#include "../../include/orbx.h"

#include <cuda_runtime.h>

#include <string>
#include <vector>

struct orbx_pipe {
    struct Slot {
        orbx_extractor *h = nullptr;
        cudaStream_t st = nullptr;
        cudaEvent_t evIn = nullptr, evDone = nullptr;
        int ticket = 0;                 // ticket of the submission this slot holds (0: none)
    };
    std::vector<Slot> slots;
    int device = 0, next = 1;
    std::string err;
};

namespace {
int pfail(orbx_pipe *p, int code, const std::string &msg) { if (p) p->err = msg; return code; }
#define PCK(x)                                                                                         \
    do {                                                                                               \
        cudaError_t e_ = (x);                                                                          \
        if (e_ != cudaSuccess) return pfail(p, ORBX_ERR_CUDA, std::string(#x ": ") + cudaGetErrorString(e_)); \
    } while (0)
} // namespace

extern "C" {

int orbx_pipe_create(const orbx_config *cfg, int depth, orbx_pipe **out)
{
    if (!cfg || !out || depth < 1 || depth > 8) return ORBX_ERR_ARG;
    orbx_pipe *p = new (std::nothrow) orbx_pipe();
    if (!p) return ORBX_ERR_NOMEM;
    *out = p;
    p->device = cfg->device;
    p->slots.resize((size_t)depth);
    for (orbx_pipe::Slot &s : p->slots) {
        const int rc = orbx_create(cfg, &s.h);
        if (rc != ORBX_OK) return pfail(p, rc, s.h ? orbx_last_error(s.h) : "invalid configuration");
        if (depth > 1) orbx_set_device_split(s.h, 1);      // the other batches in flight provide the overlap; whole-batch launches
        PCK(cudaSetDevice(p->device));
        PCK(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
        PCK(cudaEventCreateWithFlags(&s.evIn, cudaEventDisableTiming));
        PCK(cudaEventCreateWithFlags(&s.evDone, cudaEventDisableTiming));
    }
    return ORBX_OK;
}

void orbx_pipe_destroy(orbx_pipe *p)
{
    if (!p) return;
    cudaSetDevice(p->device);
    for (orbx_pipe::Slot &s : p->slots) {
        if (s.st) cudaStreamSynchronize(s.st);
        if (s.h) orbx_destroy(s.h);
        if (s.evIn) cudaEventDestroy(s.evIn);
        if (s.evDone) cudaEventDestroy(s.evDone);
        if (s.st) cudaStreamDestroy(s.st);
    }
    delete p;
}

const char *orbx_pipe_last_error(const orbx_pipe *p) { return p ? p->err.c_str() : "null handle"; }
int orbx_pipe_depth(const orbx_pipe *p) { return p ? (int)p->slots.size() : ORBX_ERR_ARG; }

int orbx_pipe_submit(orbx_pipe *p, const uint8_t *d_imgs, size_t frame_stride, size_t pitch, int batch, int width, int height,
                     void *stream, int *ticket)
{
    if (!p) return ORBX_ERR_ARG;
    if (!ticket) return pfail(p, ORBX_ERR_ARG, "null ticket");
    PCK(cudaSetDevice(p->device));
    const int id = p->next;
    orbx_pipe::Slot &s = p->slots[(size_t)id % p->slots.size()];
    // the frames are ready at the caller's position in `stream`; whatever the caller enqueued there on the slot's previous
    // results is ahead of that position too, so the slot's arenas are free once the slot's stream has passed this event
    PCK(cudaEventRecord(s.evIn, (cudaStream_t)stream));
    PCK(cudaStreamWaitEvent(s.st, s.evIn, 0));
    const int rc = orbx_extract_batch_device(s.h, d_imgs, frame_stride, pitch, batch, width, height, s.st);
    if (rc != ORBX_OK) return pfail(p, rc, orbx_last_error(s.h));
    PCK(cudaEventRecord(s.evDone, s.st));
    s.ticket = id;
    p->next = id + 1;
    *ticket = id;
    return ORBX_OK;
}

int orbx_pipe_join(orbx_pipe *p, int ticket, void *stream, const orbx_keypoint **d_kps, const uint8_t **d_desc, const int **d_counts,
                   int *kp_stride)
{
    if (!p) return ORBX_ERR_ARG;
    if (ticket < 1) return pfail(p, ORBX_ERR_ARG, "no such ticket");
    orbx_pipe::Slot &s = p->slots[(size_t)ticket % p->slots.size()];
    if (s.ticket != ticket) return pfail(p, ORBX_ERR_ARG, "the ticket's slot has been submitted to again (or never was)");
    PCK(cudaSetDevice(p->device));
    PCK(cudaStreamWaitEvent((cudaStream_t)stream, s.evDone, 0));
    const int rc = orbx_device_results(s.h, d_kps, d_desc, d_counts, kp_stride);
    if (rc != ORBX_OK) return pfail(p, rc, orbx_last_error(s.h));
    return ORBX_OK;
}

orbx_extractor *orbx_pipe_handle(orbx_pipe *p, int ticket)
{
    if (!p || ticket < 1) return nullptr;
    orbx_pipe::Slot &s = p->slots[(size_t)ticket % p->slots.size()];
    return s.ticket == ticket ? s.h : nullptr;
}

} // extern "C"
