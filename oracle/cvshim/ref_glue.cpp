// ref_glue.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Links the reference's own src/orbextractor.cpp (compiled UNMODIFIED from /root/reference by
// oracle/Makefile target `ref`) to the cv:: primitives restated in oracle/orb_oracle.c and exposes
// a C entry point, so the reference's cell loop, std::list octree, orientation and descriptor
// code run exactly as written.
//
// DistributeOctTree sorts (size, node pointer) pairs (orbextractor.cpp:825), i.e. equal-size
// ties are broken by heap address.  "canonical" mode runs the call under a monotone bump
// allocator (addresses only grow, nothing is reused) so that higher address == created later;
// "stock" mode leaves glibc malloc in charge to show the reference's own nondeterminism.
#include "orbextractor.hpp"

#include <atomic>
#include <sys/mman.h>

#include <cstdio>
#include <cstdlib>
#include <new>
#include <sstream>

extern "C" {
#include "orb_oracle.h"
}

// ---------------------------------------------------------------- cv:: primitives -> C oracle
namespace cv {
void resize(const Mat &src, Mat &dst, Size dsize, double, double, int)
{
    if (dst.rows != dsize.height || dst.cols != dsize.width) dst.create(dsize.height, dsize.width, CV_8UC1);
    orbo_resize_linear_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step);
}
void copyMakeBorder(const Mat &src, Mat &dst, int top, int bottom, int left, int right, int)
{
    if (dst.rows != src.rows + top + bottom || dst.cols != src.cols + left + right)
        dst.create(src.rows + top + bottom, src.cols + left + right, CV_8UC1);
    orbo_border_reflect101_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.step, top, bottom, left, right);
}
void FAST(const Mat &image, std::vector<KeyPoint> &keypoints, int threshold, bool)
{
    keypoints.clear();
    const int cap = image.rows * image.cols;
    if (cap <= 0) return;
    std::vector<int> xs(cap), ys(cap), sc(cap);
    const int n = orbo_fast9_nms(image.data, image.cols, image.rows, image.step, threshold, xs.data(), ys.data(), sc.data(), cap);
    for (int i = 0; i < n; i++) keypoints.push_back(KeyPoint((float)xs[i], (float)ys[i], 7.f, -1.f, (float)sc[i], 0, -1));
}
static int g_taps[7] = {18, 34, 48, 56, 48, 34, 18};
void GaussianBlur(const Mat &src, Mat &dst, Size, double, double, int)
{
    // in place in the reference (orbextractor.cpp:622): the oracle buffers the horizontal pass first
    if (dst.rows != src.rows || dst.cols != src.cols) dst.create(src.rows, src.cols, CV_8UC1);
    orbo_gaussian7_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.step, g_taps);
}
float fastAtan2(float y, float x) { return orbo_fast_atan2(y, x); }
} // namespace cv

// ---------------------------------------------------------------- monotone bump allocator
static char *g_arena = nullptr;
static size_t g_arena_size = 0;
static std::atomic<size_t> g_arena_used{0};   // atomic: the reference's OrbFrame extracts left and right in two threads
static bool g_bump_on = false;

static void arena_init()
{
    if (g_arena) return;
    g_arena_size = (size_t)16 << 30; // virtual reservation; pages are committed lazily
    void *p = mmap(nullptr, g_arena_size, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) { perror("mmap arena"); abort(); }
    g_arena = (char *)p;
}
static inline bool in_arena(void *p) { return g_arena && (char *)p >= g_arena && (char *)p < g_arena + g_arena_size; }

void *operator new(size_t n)
{
    if (g_bump_on) {
        const size_t a = g_arena_used.fetch_add((n + 15) & ~(size_t)15);   // later allocation = higher address, in every thread
        if (a + n > g_arena_size) { fprintf(stderr, "bump arena exhausted\n"); abort(); }
        return g_arena + a;
    }
    void *p = malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void *operator new[](size_t n) { return operator new(n); }
void operator delete(void *p) noexcept { if (p && !in_arena(p)) free(p); }
void operator delete[](void *p) noexcept { operator delete(p); }
void operator delete(void *p, size_t) noexcept { operator delete(p); }
void operator delete[](void *p, size_t) noexcept { operator delete(p); }

// ---------------------------------------------------------------- C entry points
extern "C" {

// canonical heap order for code outside this file (frame_glue.cpp): on -> fresh bump arena, off -> malloc again
void orbref_canonical(int on) { if (on) { arena_init(); g_arena_used = 0; } g_bump_on = on != 0; }

struct orbref_cfg { int nfeatures; float scale; int nlevels, ini_th, min_th; };

void orbref_set_taps(const int *t) { for (int k = 0; k < 7; k++) cv::g_taps[k] = t[k]; }

// One ExtractFeatures call on a fresh OrbExtractor.  canonical != 0 -> bump allocator.
// level_out (optional): nlevels pointers to caller buffers receiving each pyramid ROI tightly packed.
int orbref_extract(const orbref_cfg *c, int canonical, const uint8_t *img, int w, int h, size_t stride,
                   orbo_keypoint *kps, uint8_t *desc, int cap, uint8_t **level_out, int *level_w, int *level_h)
{
    int n = -1;
    std::streambuf *old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf()); // the constructor prints "Making extractor" (orbextractor.cpp:491)
    if (canonical) { arena_init(); g_arena_used = 0; g_bump_on = true; }
    {
        OrbExtractor ex(c->nfeatures, c->scale, c->nlevels, c->ini_th, c->min_th);
        cv::Mat image(h, w, CV_8UC1, (void *)img, stride);
        std::vector<cv::KeyPoint> keys;
        cv::Mat d;
        ex.ExtractFeatures(image, keys, d);
        n = (int)keys.size();
        if (n <= cap) {
            for (int i = 0; i < n; i++) {
                memcpy(&kps[i], &keys[i], sizeof(orbo_keypoint));
                memcpy(desc + (size_t)i * 32, d.ptr(i), 32);
            }
        } else {
            n = -1;
        }
        if (level_out)
            for (int l = 0; l < c->nlevels; l++) {
                const cv::Mat &m = ex.m_vImagePyramid[l];
                level_w[l] = m.cols; level_h[l] = m.rows;
                if (level_out[l])
                    for (int y = 0; y < m.rows; y++) memcpy(level_out[l] + (size_t)y * m.cols, m.ptr(y), m.cols);
            }
    } // every arena object dies here
    if (canonical) { g_bump_on = false; g_arena_used = 0; }
    std::cout.rdbuf(old);
    return n;
}

// persistent instance for timing the reference as it runs in production (stock malloc)
void *orbref_create(const orbref_cfg *c)
{
    std::streambuf *old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    OrbExtractor *ex = new OrbExtractor(c->nfeatures, c->scale, c->nlevels, c->ini_th, c->min_th);
    std::cout.rdbuf(old);
    return ex;
}
void orbref_destroy(void *p) { delete (OrbExtractor *)p; }
int orbref_run(void *p, const uint8_t *img, int w, int h, size_t stride, orbo_keypoint *kps, uint8_t *desc, int cap)
{
    OrbExtractor *ex = (OrbExtractor *)p;
    cv::Mat image(h, w, CV_8UC1, (void *)img, stride);
    std::vector<cv::KeyPoint> keys;
    cv::Mat d;
    ex->ExtractFeatures(image, keys, d);
    const int n = (int)keys.size();
    if (n > cap) return -1;
    for (int i = 0; i < n; i++) {
        memcpy(&kps[i], &keys[i], sizeof(orbo_keypoint));
        memcpy(desc + (size_t)i * 32, d.ptr(i), 32);
    }
    return n;
}

} // extern "C"
