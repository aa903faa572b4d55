// orbx_multi.cu -- several GPUs behind ONE extractor handle, inside the library ("orbx_extract_batch
// frame-sharded across the handle's GPUs"; 8(e): frames are independent units, no data-path collective).
//
// A C++ host of the reference -- Tracking builds its extractors once (tracking.cpp:121-127) and OrbFrame calls them per
// image (orbframe.cpp:73-76) -- reaches every GPU of the box through this handle: device slot g owns an ordinary
// orbx_extractor on its GPU plus one host thread, and a batch of B frames is cut into contiguous blocks, slot g taking
// frames [B*g/G, B*(g+1)/G).  The blocks are submitted by the slots' threads side by side (orbx_extract_batch_async on each
// GPU: H2D copies, kernels and D2H copies of all GPUs run concurrently) and the results land in disjoint slices of the
// caller's arrays, exactly where a single-GPU call would have put them.  Built on the public C ABI only.
#include "../../include/orbx.h"

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Job {
    int kind = 0;                       // 1 submit, 2 wait
    const uint8_t *const *imgs = nullptr;
    int nf = 0, width = 0, height = 0, kp_cap = 0;
    size_t pitch = 0;
    orbx_keypoint *kps = nullptr;
    uint8_t *desc = nullptr;
    int *n_out = nullptr;
    int ticket = 0;
};

struct Slot {
    orbx_extractor *h = nullptr;
    int device = 0;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    bool pending = false, done = false, quit = false;
    Job job;
    int rc = ORBX_OK;
    int subTicket[2] = {0, 0};          // the slot's own ticket of the multi-ticket with that parity (0: the slot had no frames)
};

void run(Slot *s)
{
    for (;;) {
        std::unique_lock<std::mutex> lk(s->mu);
        s->cv.wait(lk, [s] { return s->pending || s->quit; });
        if (s->quit) return;
        const Job j = s->job;
        s->pending = false;
        lk.unlock();
        int rc = ORBX_OK;
        if (j.kind == 1) {
            int t = 0;
            rc = orbx_extract_batch_async(s->h, j.imgs, j.nf, j.width, j.height, j.pitch, j.kps, j.kp_cap, j.desc, j.n_out, &t);
            s->subTicket[j.ticket & 1] = rc == ORBX_OK ? t : 0;
        } else if (j.kind == 2) {
            const int t = s->subTicket[j.ticket & 1];
            s->subTicket[j.ticket & 1] = 0;
            if (t) rc = orbx_wait(s->h, t);
        }
        lk.lock();
        s->rc = rc;
        s->done = true;
        lk.unlock();
        s->cv.notify_all();
    }
}

} // namespace

struct orbx_multi {
    std::vector<Slot *> slots;
    int maxBatch = 0, nextTicket = 1;
    bool open[2] = {false, false};
    int openId[2] = {0, 0};
    std::string err;
};

namespace {

int mfail(orbx_multi *m, int code, const std::string &msg) { if (m) m->err = msg; return code; }

void post(Slot *s, const Job &j)
{
    {
        std::lock_guard<std::mutex> lk(s->mu);
        s->job = j; s->pending = true; s->done = false;
    }
    s->cv.notify_all();
}

int collect(orbx_multi *m, const std::vector<int> &posted)
{
    int worst = ORBX_OK;
    for (int g : posted) {
        Slot *s = m->slots[g];
        std::unique_lock<std::mutex> lk(s->mu);
        s->cv.wait(lk, [s] { return s->done; });
        if (s->rc != ORBX_OK && (worst == ORBX_OK || worst == ORBX_ERR_CAPACITY)) {
            worst = s->rc;
            m->err = "device " + std::to_string(s->device) + ": " + orbx_last_error(s->h);
        }
    }
    return worst;
}

int waitTicket(orbx_multi *m, int ticket)
{
    const int p = ticket & 1;
    m->open[p] = false;
    Job j; j.kind = 2; j.ticket = ticket;
    std::vector<int> posted;
    for (size_t g = 0; g < m->slots.size(); g++) { post(m->slots[g], j); posted.push_back((int)g); }
    return collect(m, posted);
}

} // namespace

extern "C" {

int orbx_multi_create(const orbx_config *cfg, const int *devices, int n_devices, orbx_multi **out)
{
    if (!cfg || !devices || !out || n_devices < 1 || n_devices > 64 || cfg->max_batch < 1) return ORBX_ERR_ARG;
    orbx_multi *m = new (std::nothrow) orbx_multi();
    if (!m) return ORBX_ERR_NOMEM;
    *out = m;
    m->maxBatch = cfg->max_batch;
    for (int g = 0; g < n_devices; g++) {
        orbx_config c = *cfg;
        c.device = devices[g];
        c.max_batch = (cfg->max_batch + n_devices - 1) / n_devices;      // the largest block a slot can be handed
        Slot *s = new Slot();
        s->device = devices[g];
        m->slots.push_back(s);
        const int rc = orbx_create(&c, &s->h);
        if (rc != ORBX_OK) return mfail(m, rc, "device " + std::to_string(devices[g]) + ": " + (s->h ? orbx_last_error(s->h) : "out of memory"));
    }
    for (Slot *s : m->slots) s->th = std::thread(run, s);
    return ORBX_OK;
}

void orbx_multi_destroy(orbx_multi *m)
{
    if (!m) return;
    for (int p = 0; p < 2; p++) if (m->open[p]) waitTicket(m, m->openId[p]);
    for (Slot *s : m->slots) {
        if (s->th.joinable()) {
            { std::lock_guard<std::mutex> lk(s->mu); s->quit = true; }
            s->cv.notify_all();
            s->th.join();
        }
        if (s->h) orbx_destroy(s->h);
        delete s;
    }
    delete m;
}

const char *orbx_multi_last_error(const orbx_multi *m) { return m ? m->err.c_str() : "null handle"; }
int orbx_multi_devices(const orbx_multi *m) { return m ? (int)m->slots.size() : ORBX_ERR_ARG; }
int orbx_multi_max_keypoints(const orbx_multi *m) { return m && !m->slots.empty() ? orbx_max_keypoints(m->slots[0]->h) : ORBX_ERR_ARG; }
orbx_extractor *orbx_multi_handle(orbx_multi *m, int slot) { return m && slot >= 0 && slot < (int)m->slots.size() ? m->slots[slot]->h : nullptr; }

int orbx_multi_frame_range(const orbx_multi *m, int batch, int slot, int *first, int *count)
{
    if (!m || batch < 1 || slot < 0 || slot >= (int)m->slots.size()) return ORBX_ERR_ARG;
    const int G = (int)m->slots.size();
    const int lo = (int)((long long)batch * slot / G), hi = (int)((long long)batch * (slot + 1) / G);
    if (first) *first = lo;
    if (count) *count = hi - lo;
    return ORBX_OK;
}

int orbx_multi_extract_batch_async(orbx_multi *m, const uint8_t *const *imgs, int batch, int width, int height, size_t pitch,
                                   orbx_keypoint *kps, int kp_cap, uint8_t *desc, int *n_out, int *ticket)
{
    if (!m) return ORBX_ERR_ARG;
    if (!imgs || !kps || !desc || !n_out || !ticket || batch < 1 || batch > m->maxBatch || kp_cap < 1) return mfail(m, ORBX_ERR_ARG, "bad argument");
    const int id = m->nextTicket, p = id & 1;
    if (m->open[p]) {                                   // at most two calls in flight: complete the older one
        const int rc = waitTicket(m, m->openId[p]);
        if (rc != ORBX_OK && rc != ORBX_ERR_CAPACITY) return rc;
    }
    const int G = (int)m->slots.size();
    std::vector<int> posted;
    for (int g = 0; g < G; g++) {
        const int lo = (int)((long long)batch * g / G), hi = (int)((long long)batch * (g + 1) / G);
        m->slots[g]->subTicket[p] = 0;
        if (hi <= lo) continue;
        Job j; j.kind = 1; j.ticket = id;
        j.imgs = imgs + lo; j.nf = hi - lo; j.width = width; j.height = height; j.pitch = pitch; j.kp_cap = kp_cap;
        j.kps = kps + (size_t)lo * kp_cap; j.desc = desc + (size_t)lo * kp_cap * 32; j.n_out = n_out + lo;
        post(m->slots[g], j);
        posted.push_back(g);
    }
    const int rc = collect(m, posted);
    m->open[p] = true; m->openId[p] = id;
    m->nextTicket = id + 1;
    *ticket = id;
    if (rc != ORBX_OK) { waitTicket(m, id); return rc; }       // a slot refused its block: drain the others, report
    return ORBX_OK;
}

int orbx_multi_wait(orbx_multi *m, int ticket)
{
    if (!m) return ORBX_ERR_ARG;
    const int p = ticket & 1;
    if (ticket < 1 || !m->open[p] || m->openId[p] != ticket) return mfail(m, ORBX_ERR_ARG, "unknown ticket, or a ticket that has been waited for already");
    return waitTicket(m, ticket);
}

int orbx_multi_extract_batch(orbx_multi *m, const uint8_t *const *imgs, int batch, int width, int height, size_t pitch,
                             orbx_keypoint *kps, int kp_cap, uint8_t *desc, int *n_out)
{
    if (!m) return ORBX_ERR_ARG;
    for (int p = 0; p < 2; p++)
        if (m->open[p]) {
            const int rc = waitTicket(m, m->openId[p]);
            if (rc != ORBX_OK && rc != ORBX_ERR_CAPACITY) return rc;
        }
    int t = 0;
    const int rc = orbx_multi_extract_batch_async(m, imgs, batch, width, height, pitch, kps, kp_cap, desc, n_out, &t);
    if (rc != ORBX_OK) return rc;
    return orbx_multi_wait(m, t);
}

} // extern "C"
