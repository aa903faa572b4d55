// orbx_api.cu -- C ABI of the extractor (include/orbx.h): geometry, arenas, launch sequence.
//
// Host-side restatement of OrbExtractor's constructor tables (orbextractor.cpp:476-548) and of
// the per-level geometry of ComputePyramid / ComputeKeyPointsOctTree / DistributeOctTree
// (orbextractor.cpp:654-699, :906-947) with the reference's exact float32 / integer expressions.
// No pixel arithmetic happens on the host: there is no CPU fallback in this library.
#include "../../include/orbx.h"
#include "orbx_internal.h"
#include "orbx_kernels.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

namespace {

inline int cvRoundF(float v) { return (int)lrintf(v); }   // SSE cvtss2si: round half to even
inline int cvRoundD(double v) { return (int)lrint(v); }
inline int cvFloorD(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeilD(double v) { int i = (int)v; return i + (i < v); }
inline int alignUp(int v, int a) { return (v + a - 1) / a * a; }

template <typename T> struct DevBuf {
    T *p = nullptr; size_t n = 0;
    cudaError_t ensure(size_t count) {
        if (count <= n) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};
template <typename T> struct PinBuf {
    T *p = nullptr; size_t n = 0;
    cudaError_t ensure(size_t count) {
        if (count <= n) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; n = 0;
        cudaError_t e = cudaMallocHost((void **)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; n = 0; }
};

} // namespace

#define ORBX_LANES 4
struct orbx_extractor {
    orbx_config cfg;
    // constructor tables
    float sf[ORBX_MAXL], invSf[ORBX_MAXL], sigma2[ORBX_MAXL], invSigma2[ORBX_MAXL];
    int quota[ORBX_MAXL];
    int umax[16];
    int taps[7];
    // geometry of the current image size
    int curW = 0, curH = 0;
    OrbxLayout L;
    std::vector<OrbxSeg> segs;
    std::vector<OrbxRTab> rtab;
    std::vector<OrbxTile> tiles;
    int maxRows = 0, maxNodes = 0, pow2Nodes = 0;
    int deviceSplit = 0;             // parts an orbx_extract_batch_device call is cut into (0: two for batches of 16 frames and more)
    int fastWinRows = 0, fastListCap = 0;   // shared-memory geometry of k_fast_segs
    bool geomUploaded = false;
    // device state
    cudaStream_t stream = nullptr;
    // Two lanes: the frames of a call are cut in two halves (or its chunks alternate) that run side by side,
    // so the latency-bound stretches of one half (octree, kernel tails) are filled by the other half's work.
    // A lane is a main stream (lane 0: `stream`), a side stream (level-0 FAST, blur) and its fork/join events.
    struct Lane {
        cudaStream_t main = nullptr, side = nullptr;
        cudaEvent_t evFork = nullptr, evJoin = nullptr, evFast0 = nullptr, evPyr = nullptr, evStart = nullptr, evDone = nullptr;
    } lane[ORBX_LANES];
    cudaStream_t streamIn = nullptr, streamOut = nullptr;
    // orbx_extract_batch_device enqueues on the CALLER's stream and does not synchronise: evLast marks the end of that work,
    // and every later entry point orders itself behind it (stream wait, or a host wait where it is about to touch host-side
    // state the kernels read) before it reads the results or reuses the arenas.
    cudaEvent_t evLast = nullptr;
    bool lastPending = false;
    // asynchronous host calls (orbx_extract_batch_async / orbx_wait): up to two tickets in flight, slot = ticket & 1
    struct Ticket {
        bool active = false, directOut = false;
        int id = 0, batch = 0, kp_cap = 0, kpStride = 0;
        orbx_keypoint *kps = nullptr; uint8_t *desc = nullptr; int *n_out = nullptr;
        std::vector<int> cb;           // chunk boundaries of the call
    } tk[2];
    int nextTicket = 1, asyncParity = -1;
    bool inSyncCall = false;           // orbx_extract_batch in progress (chunk size heuristic)
    unsigned chunkSeq = 0;             // chunks rotate over the lanes across calls
    std::vector<int> prevCb;           // chunk boundaries of the last submitted call (its events: parity asyncParity)
    std::vector<cudaEvent_t> evIn[2], evK[2], evOut[2];   // per ticket parity and chunk: H2D done, kernels done, D2H done
    PinBuf<uint8_t> hInT[2], hDescT[2];                   // host staging for callers without pinned buffers, per parity
    PinBuf<orbx_keypoint_pod> hKpsT[2];
    PinBuf<int> hCountsT[2];
    // CUDA graphs of the per-chunk kernel pipeline of the host entry point (level-0 copy .. describe, both
    // streams of a lane): one cudaGraphLaunch replaces ~25 launches / event calls per chunk.  Keyed by
    // (first frame, frames, lane); dropped whenever geometry or an arena pointer changes.
    struct ChunkGraph { int f0, nf, lane, launches; cudaGraphExec_t exec; };
    std::vector<ChunkGraph> graphs;
    uint64_t graphSig = 0;
    OrbxTensorMaps tmaps;            // TMA descriptors of the pyramid levels (source of k_blur), host copy
    OrbxTensorMaps tmapsFast;        // same levels, box = FAST window (256 bytes x hCell+6 rows)
    OrbxTensorMaps tmapsResize;      // same levels, box = source region of a k_resize tile (192 x 48)
    OrbxTensorMaps tmapsDescA;       // same levels, box = unblurred IC_Angle patch of k_describe (48 x 31)
    OrbxTensorMaps tmapsDescB;       // levels of the BLUR slab, box = rBRIEF patch of k_describe (64 x 37)
    unsigned tmapGen = 0;            // bumped whenever the maps are re-encoded (frame size, arena pointers, frame count)
    const uint8_t *tmapBase = nullptr, *tmapBaseBlur = nullptr; int tmapFrames = 0, tmapW = 0, tmapH = 0;
    DevBuf<uint8_t> dIn;
    DevBuf<uint8_t> dPyrRaw, dBlurRaw, dDesc;
    struct { uint8_t *p = nullptr; } dPyr, dBlur;    // slab bases inside the padded allocations
    DevBuf<uint32_t> dCnt;
    DevBuf<unsigned long long> dBest;
    DevBuf<int2> dSlots;
    DevBuf<int> dLvlCount, dCounts, dDbgCount;
    DevBuf<orbx_keypoint_pod> dKps;
    DevBuf<OrbxSeg> dSegs;
    DevBuf<OrbxRTab> dRtab;
    DevBuf<OrbxTile> dTiles;
    DevBuf<OrbxDbgCand> dDbg;
    DevBuf<float> dStereo;          // uRight | depth of the last orbx_stereo_match
    DevBuf<int> dStereoI;           // sad per left keypoint | match count
    PinBuf<float> hStereo;
    PinBuf<uint8_t> hDesc, hLevel;
    PinBuf<orbx_keypoint_pod> hKps;
    PinBuf<int> hCounts;
    int dbgEnabled = 0, dbgCap = 0;
    int lastBatch = 0;
    int lastLaunches = 0;            // kernel launches enqueued by the last extract call
    int nSM = 148;
    std::string err;
};

namespace {

int fail(orbx_extractor *h, int code, const std::string &msg)
{
    if (h) h->err = msg;
    return code;
}
int failCuda(orbx_extractor *h, cudaError_t e, const char *where)
{
    return fail(h, ORBX_ERR_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}
#define CK(call)                                                      \
    do {                                                              \
        cudaError_t e_ = (call);                                      \
        if (e_ != cudaSuccess) return failCuda(h, e_, #call);         \
    } while (0)

#define CKM(call, what)                                               \
    do {                                                              \
        cudaError_t e_ = (call);                                      \
        if (e_ != cudaSuccess) return failCuda(h, e_, what);          \
    } while (0)

// order `st` behind the last device-resident extraction (a no-op when that call used `st` itself or has been waited for)
cudaError_t orderAfterLast(orbx_extractor *h, cudaStream_t st)
{
    // the last asynchronous host call: its kernels AND its D2H copies (the D2H stream is in order: the last chunk's event)
    if (h->asyncParity >= 0 && h->prevCb.size() > 1) {
        const cudaError_t e = cudaStreamWaitEvent(st, h->evOut[h->asyncParity][h->prevCb.size() - 2], 0);
        if (e != cudaSuccess) return e;
    }
    if (!h->lastPending) return cudaSuccess;
    return cudaStreamWaitEvent(st, h->evLast, 0);
}
// host-side wait for the last device-resident extraction: before tables, tensor maps or arenas it may still read are replaced
cudaError_t drainLast(orbx_extractor *h)
{
    if (!h->lastPending) return cudaSuccess;
    cudaError_t e = cudaEventSynchronize(h->evLast);
    if (e == cudaSuccess) h->lastPending = false;
    return e;
}

// constructor tables, orbextractor.cpp:492-547
void buildTables(orbx_extractor *h)
{
    const orbx_config &c = h->cfg;
    const int nl = c.nlevels;
    h->sf[0] = 1.0f; h->sigma2[0] = 1.0f;
    for (int i = 1; i < nl; i++) {
        h->sf[i] = h->sf[i - 1] * c.scale_factor;
        h->sigma2[i] = h->sf[i] * h->sf[i];
    }
    for (int i = 0; i < nl; i++) {
        h->invSf[i] = 1.0f / h->sf[i];
        h->invSigma2[i] = 1.0f / h->sigma2[i];
    }
    float factor = 1.0f / c.scale_factor;
    float nDesired = c.nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nl));
    int sum = 0;
    for (int l = 0; l < nl - 1; l++) {
        h->quota[l] = cvRoundF(nDesired);
        sum += h->quota[l];
        nDesired *= factor;
    }
    h->quota[nl - 1] = std::max(c.nfeatures - sum, 0);
    const int HP = 15;
    int v, v0, vmax = cvFloorD(HP * sqrtf(2.f) / 2 + 1);
    int vmin = cvCeilD(HP * sqrtf(2.f) / 2);
    const double hp2 = HP * HP;
    for (v = 0; v <= vmax; ++v) h->umax[v] = cvRoundD(sqrt(hp2 - v * v));
    for (v = HP, v0 = 0; v >= vmin; --v) {
        while (h->umax[v0] == h->umax[v0 + 1]) ++v0;
        h->umax[v] = v0;
        ++v0;
    }
}

// bilinear coefficient table of one axis, SURVEY A.1.  x axis: {sx0, sx1, c0 | c1 << 16, 0}; y axis: {sy0, sy1, b0, b1}.
// Returns false when the axis is not a strict downscale in the sense k_resize relies on: source index strictly
// increasing, and the second tap either the next source pixel or weightless.
bool axisTable(int ssize, int dsize, OrbxRTab *out, bool xAxis)
{
    const double invScale = (double)dsize / ssize;
    const double scale = 1.0 / invScale;
    bool ok = true;
    for (int d = 0; d < dsize; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = cvFloorD(f);
        f -= s;
        if (s < 0) { f = 0; s = 0; }
        if (s >= ssize - 1) { f = 0; s = ssize - 1; }
        const int c0 = (int16_t)cvRoundF((1.f - f) * 2048.f), c1 = (int16_t)cvRoundF(f * 2048.f);
        out[d].a = s;
        out[d].b = std::min(s + 1, ssize - 1);
        if (xAxis) { out[d].c = (int32_t)((uint32_t)c0 | (uint32_t)c1 << 16); out[d].d = 0; }
        else { out[d].c = c0; out[d].d = c1; }
        if (c0 < 0 || c1 < 0 || (out[d].b != s + 1 && c1 != 0) || (d > 0 && s <= out[d - 1].a)) ok = false;
        if (xAxis && d >= 3 && s - out[d - 3].a > 4) ok = false;   // four adjacent outputs read at most 6 adjacent source bytes
        if (!xAxis && d >= 7 && s - out[d - 7].a > 10) ok = false; // eight adjacent output rows span at most 12 source rows
    }
    return ok;
}

// geometry for an image size; returns ORBX_OK or ORBX_ERR_SHAPE with a message
int buildGeometry(orbx_extractor *h, int w, int h0, OrbxLayout &L, std::vector<OrbxSeg> &segs,
                  std::vector<OrbxRTab> &rtab, std::vector<OrbxTile> &tiles, int &maxRows, int &maxNodes, int &winRows, int &listCap)
{
    const orbx_config &c = h->cfg;
    memset(&L, 0, sizeof(L));
    segs.clear(); rtab.clear(); tiles.clear();
    winRows = 0; listCap = 0;
    L.nlevels = c.nlevels; L.iniTh = c.ini_th_fast; L.minTh = c.min_th_fast; L.tieRule = c.tie_rule;
    long long off = 0;
    int rows = 0, slots = 0;
    maxRows = 0; maxNodes = 0;
    char msg[256];
    for (int l = 0; l < c.nlevels; l++) {
        OrbxLevel &v = L.lv[l];
        const float scale = h->invSf[l];
        v.w = cvRoundF((float)w * scale);   // orbextractor.cpp:659
        v.h = cvRoundF((float)h0 * scale);
        v.pitch = alignUp(v.w, 128);
        if (off + (long long)v.pitch * v.h > 0x7fffffffLL) return fail(h, ORBX_ERR_SHAPE, "pyramid slab exceeds 2 GiB per frame");
        v.off = (int)off;
        off += (long long)v.pitch * v.h;
        // gridded FAST geometry, :914-928
        const int maxBX = v.w - ORBX_EDGE + 3, maxBY = v.h - ORBX_EDGE + 3;
        const float width = (float)(maxBX - ORBX_MINB), height = (float)(maxBY - ORBX_MINB);
        v.nCols = (int)(width / ORBX_CELL_W);
        v.nRows = (int)(height / ORBX_CELL_W);
        if (v.nCols < 1 || v.nRows < 1) {
            snprintf(msg, sizeof msg, "level %d is %dx%d: smaller than one 30-px FAST cell (the reference divides by zero, orbextractor.cpp:924-927)", l, v.w, v.h);
            return fail(h, ORBX_ERR_SHAPE, msg);
        }
        v.wCell = (int)ceilf(width / v.nCols);
        v.cellMagic = 65536 / std::max(v.wCell, 1) + 1;
        v.hCell = (int)ceilf(height / v.nRows);
        if (v.wCell > 60 || v.hCell > 60 || (long long)v.nCols * v.nRows >= 65536) return fail(h, ORBX_ERR_SHAPE, "FAST grid outside supported range");
        v.segBase = (int)segs.size();
        v.winH = v.hCell + 6;
        // cells of one cell row whose windows are processed (:935, :944), grouped into runs of <= ORBX_SEG_W tested columns
        int nProc = 0;
        for (int j = 0; j < v.nCols; j++) if (ORBX_MINB + j * v.wCell < maxBX - 6) nProc++;
        // k_fast_segs stage 1 gives a thread one quad column and a band of at most 8 rows: a run is kept narrow enough for
        // ORBX_FAST_THREADS / ceil(hCell / 8) quad columns to cover it in one round (a wider run still works, in several rounds)
        const int bandsNeeded = (v.hCell + 7) / 8;
        const int widthCap = std::min(ORBX_SEG_W, 4 * (ORBX_FAST_THREADS / bandsNeeded - 2));
        const int perSegMax = std::max(1, std::min(8, widthCap / v.wCell));
        const int nSegRow = (nProc + perSegMax - 1) / perSegMax;
        const int perSeg = nSegRow ? (nProc + nSegRow - 1) / nSegRow : 0;
        for (int i = 0; i < v.nRows; i++) {
            const int iniY = ORBX_MINB + i * v.hCell;
            int maxY = iniY + v.hCell + 6;
            if (iniY >= maxBY - 3) continue;    // :935
            if (maxY > maxBY) maxY = maxBY;
            if (maxY - iniY - 6 <= 0) continue; // no tested row (cv::FAST on a window under 7 rows finds nothing)
            for (int j0 = 0; j0 < nProc; j0 += perSeg) {
                const int j1 = std::min(j0 + perSeg, nProc);
                const int iniX = ORBX_MINB + j0 * v.wCell;
                int maxX = ORBX_MINB + (j1 - 1) * v.wCell + v.wCell + 6;   // right end of the last cell's window, :942-946
                if (maxX > maxBX) maxX = maxBX;
                OrbxSeg sg;
                sg.x0 = (uint16_t)iniX; sg.y0 = (uint16_t)iniY;
                sg.wT = (uint16_t)(maxX - iniX - 6); sg.hT = (uint8_t)(maxY - iniY - 6);
                sg.level = (uint8_t)l;
                sg.ci = (uint16_t)i; sg.cj0 = (uint16_t)j0;
                const int B0 = iniX + 3 - ((iniX - 4) & ~15);   // shared byte of the first tested pixel (TMA box is 16-byte aligned)
                const int nQ = ((B0 + (int)sg.wT - 1) >> 2) - (B0 >> 2) + 1;
                const int nBands = std::max(1, ORBX_FAST_THREADS / nQ);   // bands of rows, k_fast_segs stage 1
                sg.mQ = (uint32_t)((1u << 20) / nQ + 1) | (uint32_t)(((int)sg.hT + nBands - 1) / nBands) << 24;
                segs.push_back(sg);
                listCap = std::max(listCap, (int)sg.wT * (int)sg.hT);
            }
        }
        v.nSegs = (int)segs.size() - v.segBase;
        winRows = std::max(winRows, v.winH);
        // DistributeOctTree geometry, :684-699
        v.W = maxBX - ORBX_MINB; v.H = maxBY - ORBX_MINB;
        if (v.H <= 0 || v.W / v.H < 1) {
            snprintf(msg, sizeof msg, "level %d is %dx%d: portrait shapes make nIni = 0 in the reference (division by zero, orbextractor.cpp:684-686)", l, v.w, v.h);
            return fail(h, ORBX_ERR_SHAPE, msg);
        }
        v.nIni = (int)round((double)(v.W / v.H));
        v.hX = v.W / v.nIni;
        // largest candidate x (relative) is W-4; the reference indexes vpIniNodes[x/hX] unchecked (:710)
        if (v.nIni > ORBX_MAX_STRIPS || (v.W - 4) / v.hX >= v.nIni || v.H > 8191 || v.W > 16383) {
            snprintf(msg, sizeof msg, "level %d is %dx%d: aspect ratio outside the supported range (nIni=%d)", l, v.w, v.h, v.nIni);
            return fail(h, ORBX_ERR_SHAPE, msg);
        }
        v.quota = h->quota[l];
        v.rowBase = rows;
        rows += v.nIni * v.H;
        maxRows = std::max(maxRows, v.nIni * v.H);
        v.slotBase = slots;
        v.slotCap = std::max(v.quota, 2 * v.nIni);   // list never exceeds max(N, first-pass size)
        slots += v.slotCap;
        // nodes are disjoint row ranges of the nIni strips (DivideNode never splits in x in this fork, A.5), so a level's node
        // list holds at most nIni * H of them however large the quota is
        const int nodeCap = std::min(v.slotCap, v.nIni * v.H + 2 * v.nIni) + 2;
        if (nodeCap > ORBX_MAX_NODES) return fail(h, ORBX_ERR_SHAPE, "more than 65534 octree nodes on one level");
        maxNodes = std::max(maxNodes, nodeCap);
        v.sf = h->sf[l]; v.invSf = h->invSf[l];
        v.kpSize = 31 * (int)h->sf[l];   // :978 int cast before the multiply
        // blur tiles: 32 words x (4 bands of ORBX_BLUR_ROWS rows)
        for (int ty = 0; ty < v.h; ty += 4 * ORBX_BLUR_ROWS)
            for (int tx = 0; tx * 4 < v.w; tx += 32) {
                OrbxTile t; t.x0 = (uint16_t)tx; t.y0 = (uint16_t)ty; t.level = (uint8_t)l; t.pad[0] = t.pad[1] = t.pad[2] = 0;
                tiles.push_back(t);
            }
        if (l > 0) {
            const OrbxLevel &p = L.lv[l - 1];
            v.xtabOff = (int)rtab.size();
            rtab.resize(rtab.size() + v.w);
            const bool okX = axisTable(p.w, v.w, &rtab[v.xtabOff], true);
            v.ytabOff = (int)rtab.size();
            rtab.resize(rtab.size() + v.h);
            const bool okY = axisTable(p.h, v.h, &rtab[v.ytabOff], false);
            if (!okX || !okY) {
                snprintf(msg, sizeof msg, "level %d (%dx%d from %dx%d) is not a strict downscale by at most 1.35", l, v.w, v.h, p.w, p.h);
                return fail(h, ORBX_ERR_SHAPE, msg);
            }
        }
    }
    L.slab = (off + 255) / 256 * 256;
    L.rowsPerFrame = rows;
    L.slotsPerFrame = slots;
    L.kpStride = slots;
    L.totalSegs = (int)segs.size();
    listCap = (listCap + 7) & ~7;
    return ORBX_OK;
}

int buildTensorMaps(orbx_extractor *h, int frames);

int ensureArenas(orbx_extractor *h, int batch)
{
    const OrbxLayout &L = h->L;
    // 256-byte pads in front of and behind the slabs: edge threads may read one word outside a row
    CK(h->dPyrRaw.ensure((size_t)L.slab * batch + 512));
    CK(h->dBlurRaw.ensure((size_t)L.slab * batch + 512));
    h->dPyr.p = h->dPyrRaw.p + 256;
    h->dBlur.p = h->dBlurRaw.p + 256;
    CK(h->dCnt.ensure((size_t)L.rowsPerFrame * batch));
    CK(h->dBest.ensure((size_t)L.rowsPerFrame * batch));
    CK(h->dSlots.ensure((size_t)L.slotsPerFrame * batch));
    CK(h->dLvlCount.ensure((size_t)L.nlevels * batch));
    CK(h->dCounts.ensure((size_t)batch));
    CK(h->dKps.ensure((size_t)L.kpStride * batch));
    CK(h->dDesc.ensure((size_t)L.kpStride * batch * 32));
    if (h->dbgEnabled) {
        h->dbgCap = 0;
        for (int l = 0; l < L.nlevels; l++) h->dbgCap = std::max(h->dbgCap, (L.lv[l].w * L.lv[l].h) / 4 + 64);
        CK(h->dDbg.ensure((size_t)h->dbgCap * L.nlevels * batch));
        CK(h->dDbgCount.ensure((size_t)L.nlevels * batch));
    }
    return buildTensorMaps(h, (int)((h->dPyrRaw.n - 512) / (size_t)L.slab));
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int buildTensorMaps(orbx_extractor *h, int frames)
{
    const OrbxLayout &L = h->L;
    if (h->tmapBase == h->dPyr.p && h->tmapBaseBlur == h->dBlur.p && h->tmapFrames >= frames && h->tmapW == h->curW && h->tmapH == h->curH) return ORBX_OK;
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess) return fail(h, ORBX_ERR_CUDA, "cuTensorMapEncodeTiled not available");
        encode = (EncodeTiledFn)fn;
    }
    for (int l = 0; l < L.nlevels; l++) {
        const OrbxLevel &v = L.lv[l];
        cuuint64_t gdim[3] = {(cuuint64_t)v.w, (cuuint64_t)v.h, (cuuint64_t)frames};
        cuuint64_t gstr[2] = {(cuuint64_t)v.pitch, (cuuint64_t)L.slab};
        cuuint32_t box[3] = {160, 4 * ORBX_BLUR_ROWS + 6, 1};   // BL_BOXW x BL_BOXH of k_blur
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&h->tmaps.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)(h->dPyr.p + v.off), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        // k_fast_segs reads 32-bit elements: a box of ORBX_FAST_PITCH / 4 elements x window rows (k_fast_segs indexes its window in words; the box start is 16-byte aligned either way)
        cuuint64_t gdimF[3] = {(cuuint64_t)(v.pitch / 4), (cuuint64_t)v.h, (cuuint64_t)frames};
        cuuint32_t boxF[3] = {ORBX_FAST_PITCH / 4, (cuuint32_t)v.winH, 1};
        CUresult r2 = encode(&h->tmapsFast.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void *)(h->dPyr.p + v.off), gdimF, gstr, boxF, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cuuint32_t boxR[3] = {192, 48, 1};                             // RS_BOXW x RS_BOXH of k_resize
        CUresult r3 = encode(&h->tmapsResize.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)(h->dPyr.p + v.off), gdim, gstr, boxR, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cuuint32_t boxA[3] = {48, 31, 1}, boxB[3] = {64, 37, 1};      // DS_PA x 31 and DS_PB x 37 of k_describe
        CUresult r4 = encode(&h->tmapsDescA.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)(h->dPyr.p + v.off), gdim, gstr, boxA, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CUresult r5 = encode(&h->tmapsDescB.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)(h->dBlur.p + v.off), gdim, gstr, boxB, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r4 != CUDA_SUCCESS) r3 = r4;
        if (r5 != CUDA_SUCCESS) r3 = r5;
        if (r != CUDA_SUCCESS || r2 != CUDA_SUCCESS || r3 != CUDA_SUCCESS) {
            char msg[96];
            snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled failed for level %d (CUresult %d / %d / %d)", l, (int)r, (int)r2, (int)r3);
            return fail(h, ORBX_ERR_CUDA, msg);
        }
    }
    // the maps travel as kernel parameters (by value): launches already enqueued keep their own copies, nothing to wait for;
    // graphs captured with the previous maps are dropped through the signature
    h->tmapGen++;
    h->tmapBase = h->dPyr.p; h->tmapBaseBlur = h->dBlur.p; h->tmapFrames = frames; h->tmapW = h->curW; h->tmapH = h->curH;
    return ORBX_OK;
}

int setGeometry(orbx_extractor *h, int w, int hh)
{
    if (w == h->curW && hh == h->curH && h->geomUploaded) return ORBX_OK;
    OrbxLayout L; std::vector<OrbxSeg> segs; std::vector<OrbxRTab> rtab; std::vector<OrbxTile> tiles; int maxRows, maxNodes, winRows, listCap;
    int rc = buildGeometry(h, w, hh, L, segs, rtab, tiles, maxRows, maxNodes, winRows, listCap);
    if (rc != ORBX_OK) return rc;
    // everything that can reject the shape is checked BEFORE the handle's geometry is touched: a refused size leaves the
    // previous one fully usable
    if (octree_smem_bytes(maxRows, maxNodes) > 200 * 1024) {
        char msg[160];
        snprintf(msg, sizeof msg, "octree needs %zu bytes of shared memory (%d nodes per level, %d strip rows): more than 200 KB -- fewer features or a shorter level",
                 octree_smem_bytes(maxRows, maxNodes), maxNodes, maxRows);
        return fail(h, ORBX_ERR_SHAPE, msg);
    }
    // in-flight work (the handle's lanes, a caller's stream of the last device-resident call) may still read the old tables
    CK(drainLast(h));
    for (int i = 0; i < ORBX_LANES; i++) {
        CK(cudaStreamSynchronize(h->lane[i].main));
        CK(cudaStreamSynchronize(h->lane[i].side));
    }
    if (h->streamIn) CK(cudaStreamSynchronize(h->streamIn));
    if (h->streamOut) CK(cudaStreamSynchronize(h->streamOut));
    h->prevCb.clear();
    // from here on the handle describes no size until the new tables are on the device: a failure below (out of memory)
    // cannot leave the old curW / curH paired with the new layout
    h->geomUploaded = false; h->curW = 0; h->curH = 0;
    h->L = L; h->segs.swap(segs); h->rtab.swap(rtab); h->tiles.swap(tiles);
    h->maxRows = maxRows; h->maxNodes = maxNodes;
    h->fastWinRows = winRows; h->fastListCap = listCap;
    int p2 = 2; while (p2 < maxNodes) p2 <<= 1;
    h->pow2Nodes = p2;
    CK(h->dSegs.ensure(h->segs.size()));
    CK(h->dRtab.ensure(std::max<size_t>(h->rtab.size(), 1)));
    CK(h->dTiles.ensure(h->tiles.size()));
    CK(cudaMemcpyAsync(h->dSegs.p, h->segs.data(), h->segs.size() * sizeof(OrbxSeg), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->dTiles.p, h->tiles.data(), h->tiles.size() * sizeof(OrbxTile), cudaMemcpyHostToDevice, h->stream));
    if (!h->rtab.empty())
        CK(cudaMemcpyAsync(h->dRtab.p, h->rtab.data(), h->rtab.size() * sizeof(OrbxRTab), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->curW = w; h->curH = hh; h->geomUploaded = true;
    return ORBX_OK;
}

// enqueue every stage after level 0 is in place
int enqueuePipeline(orbx_extractor *h, int f0, int batch, cudaStream_t st, const orbx_extractor::Lane &ln)
{
    const OrbxLayout &L = h->L;
    uint8_t *pyr = h->dPyr.p + (size_t)f0 * L.slab;
    uint32_t *cnt = h->dCnt.p + (size_t)f0 * L.rowsPerFrame;
    unsigned long long *best = h->dBest.p + (size_t)f0 * L.rowsPerFrame;
    int2 *slots = h->dSlots.p + (size_t)f0 * L.slotsPerFrame;
    int *lvlCount = h->dLvlCount.p + (size_t)f0 * L.nlevels;
    OrbxDbgCand *dbg = nullptr; int *dbgCount = nullptr;
    if (h->dbgEnabled) {
        dbg = h->dDbg.p + (size_t)f0 * L.nlevels * h->dbgCap; dbgCount = h->dDbgCount.p + (size_t)f0 * L.nlevels;
        CK(cudaMemsetAsync(dbgCount, 0, (size_t)L.nlevels * batch * sizeof(int), st));
    }
    CK(cudaMemsetAsync(cnt, 0, (size_t)L.rowsPerFrame * batch * sizeof(uint32_t), st));
    CK(cudaMemsetAsync(best, 0, (size_t)L.rowsPerFrame * batch * sizeof(unsigned long long), st));
    // Level 0 is in place: its FAST segments (a third of all tested pixels, issue-bound) run on the side stream
    // beside the resize chain (7 dependent, latency-bound launches); the blur follows there once the
    // chain is done, beside FAST of the upper levels + octree on the main stream.
    const int segs0 = L.lv[0].nSegs;
    CK(cudaEventRecord(ln.evFork, st));
    CK(cudaStreamWaitEvent(ln.side, ln.evFork, 0));
    CK(launch_fast(h->tmapsFast, f0, L, h->dSegs.p, 0, segs0, cnt, best, dbg, dbgCount, h->dbgCap, h->fastWinRows, h->fastListCap, batch, ln.side));
    CK(cudaEventRecord(ln.evFast0, ln.side));
    for (int l = 1; l < L.nlevels; l++) {
        launch_resize(h->tmapsResize, f0, pyr, L, l, (const int4 *)h->dRtab.p, batch, h->nSM, st);
        CKM(cudaGetLastError(), "k_resize launch");
    }
    CK(cudaEventRecord(ln.evPyr, st));
    CK(cudaStreamWaitEvent(ln.side, ln.evPyr, 0));
    // FAST of levels 1.. is enqueued BEFORE the blur: CTAs are dispatched in launch order, so the blur fills FAST's tail and keeps
    // the SMs busy while the octree (a few latency-bound CTAs) runs, instead of the other way round
    CK(launch_fast(h->tmapsFast, f0, L, h->dSegs.p, segs0, L.totalSegs - segs0, cnt, best, dbg, dbgCount, h->dbgCap, h->fastWinRows, h->fastListCap, batch, st));
    CK(cudaStreamWaitEvent(st, ln.evFast0, 0));
    CK(launch_octree(L, cnt, best, slots, lvlCount, h->maxRows, h->maxNodes, batch, st));
    launch_blur(h->tmaps, h->dBlur.p, L, h->dTiles.p, (int)h->tiles.size(), h->taps, f0, batch, ln.side);
    CKM(cudaGetLastError(), "k_blur launch");
    CK(cudaEventRecord(ln.evJoin, ln.side));
    CK(cudaStreamWaitEvent(st, ln.evJoin, 0));
    launch_describe(h->tmapsDescA, h->tmapsDescB, f0, L, slots, lvlCount, h->umax, h->dKps.p + (size_t)f0 * L.kpStride,
                    h->dDesc.p + (size_t)f0 * L.kpStride * 32, h->dCounts.p + f0, batch, st);
    CKM(cudaGetLastError(), "k_describe launch");
    // level-0 copy (by the caller of this function) + resize chain + FAST (level 0 | upper levels) + blur + octree + describe
    h->lastLaunches += 1 + (L.nlevels - 1) + (segs0 > 0) + (L.totalSegs - segs0 > 0) + 1 + 1 + 1;
    return ORBX_OK;
}

uint64_t arenaSignature(const orbx_extractor *h)
{
    const void *ptrs[] = {h->dIn.p, h->dPyr.p, h->dBlur.p, h->dCnt.p, h->dBest.p, h->dSlots.p, h->dLvlCount.p, h->dCounts.p,
                          h->dKps.p, h->dDesc.p, h->dSegs.p, h->dRtab.p, h->dTiles.p, h->dDbg.p, h->dDbgCount.p};
    uint64_t x = 1469598103934665603ull;
    auto mix = [&x](uint64_t v) { x = (x ^ v) * 1099511628211ull; };
    for (const void *q : ptrs) mix((uint64_t)(uintptr_t)q);
    mix((uint64_t)h->curW); mix((uint64_t)h->curH); mix((uint64_t)h->dbgEnabled); mix((uint64_t)h->dbgCap);
    mix((uint64_t)h->tmapGen);                 // kernel parameters of a captured graph hold the maps by value
    return x;
}

void dropGraphs(orbx_extractor *h)
{
    for (auto &g : h->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    h->graphs.clear();
}

// the kernel pipeline of frames [f0, f0+nf) of the host staging buffer as an executable graph on lane `li`
int chunkGraph(orbx_extractor *h, int f0, int nf, int li, size_t frameBytes, int width, cudaGraphExec_t *out, int *launches)
{
    const uint64_t sig = arenaSignature(h);
    if (sig != h->graphSig) { dropGraphs(h); h->graphSig = sig; }
    for (auto &g : h->graphs) if (g.f0 == f0 && g.nf == nf && g.lane == li) { *out = g.exec; *launches = g.launches; return ORBX_OK; }
    const orbx_extractor::Lane &ln = h->lane[li];
    const OrbxLayout &L = h->L;
    cudaGraph_t graph = nullptr;
    const int before = h->lastLaunches;
    CK(cudaStreamBeginCapture(ln.main, cudaStreamCaptureModeRelaxed));
    launch_copy_level0(h->dIn.p + (size_t)f0 * frameBytes, frameBytes, (size_t)width, h->dPyr.p + (size_t)f0 * L.slab, L, nf, ln.main);
    int rc = cudaGetLastError() == cudaSuccess ? enqueuePipeline(h, f0, nf, ln.main, ln) : fail(h, ORBX_ERR_CUDA, "k_copy_level0 launch failed");
    cudaError_t e = cudaStreamEndCapture(ln.main, &graph);
    if (rc != ORBX_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return failCuda(h, e, "cudaStreamEndCapture");
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return failCuda(h, e, "cudaGraphInstantiate");
    *launches = h->lastLaunches - before;
    h->lastLaunches = before;
    h->graphs.push_back({f0, nf, li, *launches, exec});
    *out = exec;
    return ORBX_OK;
}

bool isPinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

} // namespace

extern "C" {

const char *orbx_version(void) { return "orbx 0.1 sm_100a"; }

int orbx_create(const orbx_config *cfg, orbx_extractor **out)
{
    if (!cfg || !out) return ORBX_ERR_ARG;
    *out = nullptr;
    if (cfg->nlevels < 1 || cfg->nlevels > ORBX_MAXL || cfg->nfeatures < 1 || !(cfg->scale_factor > 1.0f) ||
        cfg->scale_factor > 1.35f ||   /* k_resize: the source region of a 128 x 32 tile (128 s + 17 columns) must fit its 192 x 48 TMA box */
        cfg->min_th_fast < 1 || cfg->min_th_fast > 254 ||   /* 8-bit scores; k_fast_segs' packed quick reject clamps its own threshold to 127 */
        cfg->ini_th_fast < cfg->min_th_fast || cfg->ini_th_fast > 254 ||
        cfg->max_batch < 1 || cfg->max_width < 1 || cfg->max_height < 1)
        return ORBX_ERR_ARG;
    orbx_extractor *h = new (std::nothrow) orbx_extractor();
    if (!h) return ORBX_ERR_NOMEM;
    h->cfg = *cfg;
    bool zero = true;
    int tapSum = 0;
    for (int k = 0; k < 7; k++) {
        zero = zero && cfg->blur_taps[k] == 0;
        if (cfg->blur_taps[k] < 0 || cfg->blur_taps[k] > 255) return ORBX_ERR_ARG;
        tapSum += cfg->blur_taps[k];
    }
    if (tapSum > 257) return ORBX_ERR_ARG;   // horizontal sums are carried in 16 bits (255 * 257 = 65535)
    static const int kDefaultTaps[7] = {18, 34, 48, 56, 48, 34, 18};
    for (int k = 0; k < 7; k++) h->taps[k] = zero ? kDefaultTaps[k] : cfg->blur_taps[k];
    buildTables(h);
    *out = h; // returned even on failure below so the caller can read orbx_last_error, then destroy
    int devCount = 0;
    cudaError_t e = cudaGetDeviceCount(&devCount);
    if (e != cudaSuccess || devCount <= cfg->device || cfg->device < 0) {
        cudaGetLastError();
        return fail(h, ORBX_ERR_CUDA, e != cudaSuccess ? std::string("no CUDA device: ") + cudaGetErrorString(e)
                                                        : std::string("device ordinal out of range"));
    }
    CK(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) return fail(h, ORBX_ERR_CUDA, "liborbx is built for sm_100a only (no other code path exists)");
    h->nSM = prop.multiProcessorCount;
    for (int i = 0; i < ORBX_LANES; i++) {
        orbx_extractor::Lane &ln = h->lane[i];
        CK(cudaStreamCreateWithFlags(&ln.main, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ln.side, cudaStreamNonBlocking));
        cudaEvent_t *evs[6] = {&ln.evFork, &ln.evJoin, &ln.evFast0, &ln.evPyr, &ln.evStart, &ln.evDone};
        for (cudaEvent_t *e : evs) CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    }
    h->stream = h->lane[0].main;
    CK(cudaStreamCreateWithFlags(&h->streamIn, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->streamOut, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->evLast, cudaEventDisableTiming));
    // size the arenas for the declared maximum so the hot path never allocates
    int rc = setGeometry(h, cfg->max_width, cfg->max_height);
    if (rc != ORBX_OK) return rc;
    rc = ensureArenas(h, cfg->max_batch);
    if (rc != ORBX_OK) return rc;
    CK(h->hCounts.ensure((size_t)cfg->max_batch));
    for (int p = 0; p < 2; p++) CK(h->hCountsT[p].ensure((size_t)cfg->max_batch));
    return ORBX_OK;
}

void orbx_destroy(orbx_extractor *h)
{
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    for (int i = 0; i < ORBX_LANES; i++) {
        if (h->lane[i].main) cudaStreamSynchronize(h->lane[i].main);
        if (h->lane[i].side) cudaStreamSynchronize(h->lane[i].side);
    }
    if (h->streamIn) cudaStreamSynchronize(h->streamIn);
    if (h->streamOut) cudaStreamSynchronize(h->streamOut);
    if (h->evLast) { if (h->lastPending) cudaEventSynchronize(h->evLast); cudaEventDestroy(h->evLast); }
    dropGraphs(h);
    for (int p = 0; p < 2; p++) {
        for (std::vector<cudaEvent_t> *ev : {&h->evIn[p], &h->evK[p], &h->evOut[p]})
            for (cudaEvent_t e : *ev) if (e) cudaEventDestroy(e);
        h->hInT[p].release(); h->hDescT[p].release(); h->hKpsT[p].release(); h->hCountsT[p].release();
    }
    h->dIn.release();
    h->dPyrRaw.release(); h->dBlurRaw.release(); h->dDesc.release(); h->dCnt.release(); h->dBest.release();
    h->dSlots.release(); h->dLvlCount.release(); h->dCounts.release(); h->dDbgCount.release();
    h->dKps.release(); h->dSegs.release(); h->dRtab.release(); h->dTiles.release(); h->dStereo.release(); h->dStereoI.release(); h->hStereo.release(); h->dDbg.release();
    h->hDesc.release(); h->hLevel.release(); h->hKps.release(); h->hCounts.release();
    for (int i = 0; i < ORBX_LANES; i++) {
        orbx_extractor::Lane &ln = h->lane[i];
        cudaEvent_t evs[6] = {ln.evFork, ln.evJoin, ln.evFast0, ln.evPyr, ln.evStart, ln.evDone};
        for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
        if (ln.main) cudaStreamDestroy(ln.main);
        if (ln.side) cudaStreamDestroy(ln.side);
    }
    if (h->streamIn) cudaStreamDestroy(h->streamIn);
    if (h->streamOut) cudaStreamDestroy(h->streamOut);
    delete h;
}

const char *orbx_last_error(const orbx_extractor *h) { return h ? h->err.c_str() : "null handle"; }

int orbx_max_keypoints(const orbx_extractor *h) { return h ? h->L.kpStride : ORBX_ERR_ARG; }

int orbx_last_launches(const orbx_extractor *h) { return h ? h->lastLaunches : ORBX_ERR_ARG; }

int orbx_scale_tables(const orbx_extractor *h, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2, int *quota)
{
    if (!h) return ORBX_ERR_ARG;
    for (int l = 0; l < h->cfg.nlevels; l++) {
        if (scale) scale[l] = h->sf[l];
        if (inv_scale) inv_scale[l] = h->invSf[l];
        if (sigma2) sigma2[l] = h->sigma2[l];
        if (inv_sigma2) inv_sigma2[l] = h->invSigma2[l];
        if (quota) quota[l] = h->quota[l];
    }
    return h->cfg.nlevels;
}

int orbx_extract_batch_device(orbx_extractor *h, const uint8_t *d_imgs, size_t frame_stride, size_t pitch,
                              int batch, int width, int height, void *stream)
{
    if (!h) return ORBX_ERR_ARG;
    if (!d_imgs || batch < 1 || batch > h->cfg.max_batch || width < 1 || height < 1 || pitch < (size_t)width)
        return fail(h, ORBX_ERR_ARG, "bad argument");
    CK(cudaSetDevice(h->cfg.device));
    int rc = setGeometry(h, width, height);
    if (rc != ORBX_OK) return rc;
    rc = ensureArenas(h, batch);
    if (rc != ORBX_OK) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    CK(orderAfterLast(h, st));        // the arenas are reused: a previous call on another stream must have finished with them
    h->lastBatch = batch;
    h->lastLaunches = 0;
    static const int splitEnv = getenv("ORBX_SPLIT") ? atoi(getenv("ORBX_SPLIT")) : 0;
    const int nSplit = std::min(batch, splitEnv > 0 ? std::min(splitEnv, ORBX_LANES) : h->deviceSplit > 0 ? std::min(h->deviceSplit, ORBX_LANES) : (batch >= 16 ? 2 : 1));
    if (nSplit == 1) {
        launch_copy_level0(d_imgs, frame_stride, pitch, h->dPyr.p, h->L, batch, st);
        rc = enqueuePipeline(h, 0, batch, st, h->lane[0]);
        if (rc != ORBX_OK) return rc;
        CK(cudaEventRecord(h->evLast, st));
        h->lastPending = true;
        return ORBX_OK;
    }
    // nSplit parts side by side: the first on the caller's stream, the others on lanes 1.., joined back at the end
    CK(cudaEventRecord(h->lane[0].evStart, st));
    for (int k = 0; k < nSplit; k++) {
        const int b0 = (int)((long long)batch * k / nSplit), b1 = (int)((long long)batch * (k + 1) / nSplit);
        const orbx_extractor::Lane &ln = h->lane[k];
        cudaStream_t sk = k == 0 ? st : ln.main;
        if (k > 0) CK(cudaStreamWaitEvent(sk, h->lane[0].evStart, 0));
        launch_copy_level0(d_imgs + (size_t)b0 * frame_stride, frame_stride, pitch, h->dPyr.p + (size_t)b0 * h->L.slab, h->L, b1 - b0, sk);
        rc = enqueuePipeline(h, b0, b1 - b0, sk, ln);
        if (rc != ORBX_OK) return rc;
        if (k > 0) {
            CK(cudaEventRecord(ln.evDone, sk));
            CK(cudaStreamWaitEvent(st, ln.evDone, 0));
        }
    }
    CK(cudaEventRecord(h->evLast, st));
    h->lastPending = true;
    return ORBX_OK;
}

int orbx_set_device_split(orbx_extractor *h, int parts)
{
    if (!h) return ORBX_ERR_ARG;
    if (parts < 0 || parts > ORBX_LANES) return fail(h, ORBX_ERR_ARG, "parts must be 0 (default) .. 4");
    h->deviceSplit = parts;
    return ORBX_OK;
}

int orbx_device_results(orbx_extractor *h, const orbx_keypoint **d_kps, const uint8_t **d_desc,
                        const int **d_counts, int *kp_stride)
{
    if (!h) return ORBX_ERR_ARG;
    if (d_kps) *d_kps = (const orbx_keypoint *)h->dKps.p;
    if (d_desc) *d_desc = h->dDesc.p;
    if (d_counts) *d_counts = h->dCounts.p;
    if (kp_stride) *kp_stride = h->L.kpStride;
    return ORBX_OK;
}

} // extern "C"

namespace {

// Host entry points.  A call is cut into chunks that flow through three streams -- H2D of chunk c+1, kernels of chunk c
// and D2H of chunk c-1 overlap.  orbx_extract_batch_async only ENQUEUES a call and hands back a ticket; orbx_wait blocks
// until that call's results are in the caller's arrays.  Two calls may be in flight: while the tail of call k (last
// chunks' kernels, last D2H) drains, the H2D of call k+1 already runs, so the copy engine feeding the GPU never idles
// between calls (the reference has the same shape: extract in threads, consume later, orbframe.cpp:73-78).
// Both calls share ONE set of device arenas; a chunk of call k+1 depends, through events, only on the chunks of call k that
// cover the same frame slots: its H2D waits for their kernels (which consumed the staged input), its kernels wait for their
// D2H (which read the result records).  Host-side staging (callers without pinned buffers) is double-buffered by ticket parity.
int finishTicket(orbx_extractor *h, int p)
{
    orbx_extractor::Ticket &t = h->tk[p];
    if (!t.active) return ORBX_OK;
    t.active = false;
    CK(cudaSetDevice(h->cfg.device));
    int status = ORBX_OK;
    const int stride = t.kpStride;
    for (size_t c = 0; c + 1 < t.cb.size(); c++) {
        const int f0 = t.cb[c], f1 = t.cb[c + 1];
        if (f1 <= f0) continue;
        CK(cudaEventSynchronize(h->evOut[p][c]));
        for (int f = f0; f < f1; f++) {
            const int n = h->hCountsT[p].p[f];
            if (n > t.kp_cap) {
                char msg[128];
                snprintf(msg, sizeof msg, "frame %d produced %d keypoints, kp_cap is %d", f, n, t.kp_cap);
                status = fail(h, ORBX_ERR_CAPACITY, msg);
                continue;
            }
            if (!t.directOut) {
                memcpy(t.kps + (size_t)f * t.kp_cap, h->hKpsT[p].p + (size_t)f * stride, sizeof(orbx_keypoint) * n);
                memcpy(t.desc + (size_t)f * t.kp_cap * 32, h->hDescT[p].p + (size_t)f * stride * 32, (size_t)n * 32);
            }
            t.n_out[f] = n;
        }
    }
    return status;
}

} // namespace

extern "C" {

int orbx_extract_batch_async(orbx_extractor *h, const uint8_t *const *imgs, int batch, int width, int height,
                             size_t pitch, orbx_keypoint *kps, int kp_cap, uint8_t *desc, int *n_out, int *ticket)
{
    if (!h) return ORBX_ERR_ARG;
    if (!imgs || !kps || !desc || !n_out || !ticket || batch < 1 || batch > h->cfg.max_batch || width < 1 || height < 1 ||
        pitch < (size_t)width || kp_cap < 1)
        return fail(h, ORBX_ERR_ARG, "bad argument");
    for (int f = 0; f < batch; f++) if (!imgs[f]) return fail(h, ORBX_ERR_ARG, "null image pointer");
    CK(cudaSetDevice(h->cfg.device));
    const int id = h->nextTicket, p = id & 1;
    // at most two calls in flight: the call before the previous one is completed here if the caller has not waited for it
    int rc = finishTicket(h, p);
    if (rc != ORBX_OK && rc != ORBX_ERR_CAPACITY) return rc;
    CK(drainLast(h));                 // a device-resident call on a caller's stream may still be using the arenas
    if (width != h->curW || height != h->curH || !h->geomUploaded) {
        // another image size re-lays the arenas: nothing of the previous call may still be moving
        CK(cudaStreamSynchronize(h->streamIn));
        CK(cudaStreamSynchronize(h->streamOut));
        h->prevCb.clear();
    }
    rc = setGeometry(h, width, height);
    if (rc != ORBX_OK) return rc;
    rc = ensureArenas(h, batch);
    if (rc != ORBX_OK) return rc;
    const OrbxLayout &L = h->L;
    const size_t frameBytes = (size_t)width * height;
    CK(h->dIn.ensure(frameBytes * batch + 64));
    CK(h->hCountsT[p].ensure((size_t)std::max(batch, h->cfg.max_batch)));
    const bool pinnedIn = isPinned(imgs[0]) && isPinned(imgs[batch - 1]);
    if (!pinnedIn) CK(h->hInT[p].ensure(frameBytes * batch));
    // results can be DMA'd straight into the caller's arrays when those are pinned and use our stride
    const bool directOut = kp_cap == L.kpStride && isPinned(kps) && isPinned(desc);
    if (!directOut) {
        CK(h->hKpsT[p].ensure((size_t)L.kpStride * batch));
        CK(h->hDescT[p].ensure((size_t)L.kpStride * batch * 32));
    }
    orbx_keypoint_pod *hk = directOut ? (orbx_keypoint_pod *)kps : h->hKpsT[p].p;
    uint8_t *hd = directOut ? desc : h->hDescT[p].p;

    // chunk boundaries: cb[c] .. cb[c+1].  ORBX_CHUNK_PLAN="4,12,16,..." gives explicit sizes (the last chunk takes the rest),
    // ORBX_CHUNKS=n gives n equal chunks
    // default: a synchronous call (nothing behind it to cover its tail) uses chunks of about 4 MB of input (8 KITTI-sized
    // frames) so that the last chunk's kernels and D2H are short; with a second call in flight the tail is covered by that
    // call's H2D and larger chunks (about 10 MB, fewer launches and copies) are faster: 110.8 k frames/s with 3 chunks of a
    // 64-frame KITTI batch, 108.5 k with 4, 99.5 k with 8, 64.7 k with one (profiles/r2_e2e_plans.txt)
    const int chunkFrames = std::max(1, (int)((size_t)(h->inSyncCall ? (4u << 20) : (21u << 19)) / frameBytes));
    int nChunks = std::max(1, std::min(16, (batch + chunkFrames - 1) / chunkFrames));
    if (const char *e = getenv("ORBX_CHUNKS")) nChunks = std::max(1, std::min(batch, atoi(e)));
    std::vector<int> cb;
    if (const char *e = getenv("ORBX_CHUNK_PLAN")) {
        cb.push_back(0);
        for (const char *q = e; *q && cb.back() < batch;) {
            const int v = atoi(q);
            if (v > 0) cb.push_back(std::min(batch, cb.back() + v));
            while (*q && *q != ',') q++;
            if (*q == ',') q++;
        }
        if (cb.back() < batch) cb.push_back(batch);
        nChunks = (int)cb.size() - 1;
    } else {
        for (int c = 0; c <= nChunks; c++) cb.push_back((int)((long long)batch * c / nChunks));
    }
    static const int nLanes = getenv("ORBX_LANES") ? std::max(1, std::min(ORBX_LANES, atoi(getenv("ORBX_LANES")))) : 3;
    for (std::vector<cudaEvent_t> *ev : {&h->evIn[p], &h->evK[p], &h->evOut[p]})
        while ((int)ev->size() < nChunks) {
            cudaEvent_t e = nullptr;
            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ev->push_back(e);
        }
    cudaStream_t sIn = h->streamIn, sOut = h->streamOut;
    h->lastLaunches = 0;
    static const bool useGraphs = !(getenv("ORBX_NO_GRAPH") && atoi(getenv("ORBX_NO_GRAPH")));
    const int pp = p ^ 1;                                       // parity of the previous call (its events)
    const std::vector<int> &pcb = h->prevCb;
    for (int c = 0; c < nChunks; c++) {
        const int f0 = cb[c], f1 = cb[c + 1];
        const int nf = f1 - f0;
        if (nf <= 0) continue;
        // ---- H2D into the tight staging buffer, once the previous call's kernels have consumed these slots
        for (size_t j = 0; j + 1 < pcb.size(); j++)
            if (pcb[j] < f1 && pcb[j + 1] > f0 && pcb[j + 1] > pcb[j]) CK(cudaStreamWaitEvent(sIn, h->evK[pp][j], 0));
        // one copy per run of frames that are contiguous on the host
        for (int f = f0; f < f1;) {
            int g = f + 1;
            if (pitch == (size_t)width && pinnedIn)
                while (g < f1 && imgs[g] == imgs[g - 1] + frameBytes) g++;
            uint8_t *dst = h->dIn.p + (size_t)f * frameBytes;
            if (pinnedIn && pitch == (size_t)width) {
                CK(cudaMemcpyAsync(dst, imgs[f], frameBytes * (g - f), cudaMemcpyHostToDevice, sIn));
            } else if (pinnedIn) {
                CK(cudaMemcpy2DAsync(dst, width, imgs[f], pitch, width, height, cudaMemcpyHostToDevice, sIn));
            } else {
                uint8_t *stage = h->hInT[p].p + (size_t)f * frameBytes;
                for (int y = 0; y < height; y++) memcpy(stage + (size_t)y * width, imgs[f] + (size_t)y * pitch, width);
                CK(cudaMemcpyAsync(dst, stage, frameBytes, cudaMemcpyHostToDevice, sIn));
            }
            f = g;
        }
        CK(cudaEventRecord(h->evIn[p][c], sIn));
        // ---- kernels of this chunk (chunks rotate over the lanes, across calls too), once the previous call's results
        // of these slots have left for the host
        const int li = (int)(h->chunkSeq++ % (unsigned)nLanes);
        const orbx_extractor::Lane &ln = h->lane[li];
        cudaStream_t sK = ln.main;
        CK(cudaStreamWaitEvent(sK, h->evIn[p][c], 0));
        for (size_t j = 0; j + 1 < pcb.size(); j++)
            if (pcb[j] < f1 && pcb[j + 1] > f0 && pcb[j + 1] > pcb[j]) CK(cudaStreamWaitEvent(sK, h->evOut[pp][j], 0));
        if (useGraphs) {
            cudaGraphExec_t g = nullptr;
            int nl = 0;
            rc = chunkGraph(h, f0, nf, li, frameBytes, width, &g, &nl);
            if (rc != ORBX_OK) return rc;
            CK(cudaGraphLaunch(g, sK));
            h->lastLaunches += nl;
        } else {
            launch_copy_level0(h->dIn.p + (size_t)f0 * frameBytes, frameBytes, (size_t)width, h->dPyr.p + (size_t)f0 * L.slab, L, nf, sK);
            rc = enqueuePipeline(h, f0, nf, sK, ln);
            if (rc != ORBX_OK) return rc;
        }
        CK(cudaEventRecord(h->evK[p][c], sK));
        // ---- D2H of this chunk's results
        CK(cudaStreamWaitEvent(sOut, h->evK[p][c], 0));
        CK(cudaMemcpyAsync(h->hCountsT[p].p + f0, h->dCounts.p + f0, sizeof(int) * nf, cudaMemcpyDeviceToHost, sOut));
        CK(cudaMemcpyAsync(hk + (size_t)f0 * L.kpStride, h->dKps.p + (size_t)f0 * L.kpStride,
                           sizeof(orbx_keypoint_pod) * (size_t)L.kpStride * nf, cudaMemcpyDeviceToHost, sOut));
        CK(cudaMemcpyAsync(hd + (size_t)f0 * L.kpStride * 32, h->dDesc.p + (size_t)f0 * L.kpStride * 32,
                           (size_t)L.kpStride * nf * 32, cudaMemcpyDeviceToHost, sOut));
        CK(cudaEventRecord(h->evOut[p][c], sOut));
    }
    h->lastBatch = batch;
    orbx_extractor::Ticket &t = h->tk[p];
    t.active = true; t.id = id; t.batch = batch; t.kp_cap = kp_cap; t.kpStride = L.kpStride; t.directOut = directOut;
    t.kps = kps; t.desc = desc; t.n_out = n_out; t.cb = cb;
    h->prevCb = cb;
    h->asyncParity = p;
    h->nextTicket = id + 1;
    *ticket = id;
    return ORBX_OK;
}

int orbx_wait(orbx_extractor *h, int ticket)
{
    if (!h) return ORBX_ERR_ARG;
    const int p = ticket & 1;
    if (ticket < 1 || !h->tk[p].active || h->tk[p].id != ticket) return fail(h, ORBX_ERR_ARG, "unknown ticket, or a ticket that has been waited for already");
    return finishTicket(h, p);
}

int orbx_extract_batch(orbx_extractor *h, const uint8_t *const *imgs, int batch, int width, int height,
                       size_t pitch, orbx_keypoint *kps, int kp_cap, uint8_t *desc, int *n_out)
{
    if (!h) return ORBX_ERR_ARG;
    // the synchronous call: nothing else in flight when it returns (an outstanding asynchronous call is completed first)
    for (int p = 0; p < 2; p++) {
        const int rc = finishTicket(h, (h->nextTicket + p) & 1);
        if (rc != ORBX_OK && rc != ORBX_ERR_CAPACITY) return rc;
    }
    int ticket = 0;
    h->inSyncCall = true;
    const int rc = orbx_extract_batch_async(h, imgs, batch, width, height, pitch, kps, kp_cap, desc, n_out, &ticket);
    h->inSyncCall = false;
    if (rc != ORBX_OK) return rc;
    return orbx_wait(h, ticket);
}

int orbx_fetch_results(orbx_extractor *h, void *stream, orbx_keypoint *kps, int kp_cap, uint8_t *desc, int *n_out)
{
    if (!h) return ORBX_ERR_ARG;
    if (!kps || !desc || !n_out || kp_cap < 1 || h->lastBatch < 1) return fail(h, ORBX_ERR_ARG, "bad argument or no previous call");
    CK(cudaSetDevice(h->cfg.device));
    const OrbxLayout &L = h->L;
    const int batch = h->lastBatch;
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    CK(orderAfterLast(h, st));
    CK(h->hKps.ensure((size_t)L.kpStride * batch));
    CK(h->hDesc.ensure((size_t)L.kpStride * batch * 32));
    CK(h->hCounts.ensure((size_t)batch));
    CK(cudaMemcpyAsync(h->hCounts.p, h->dCounts.p, sizeof(int) * batch, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h->hKps.p, h->dKps.p, sizeof(orbx_keypoint_pod) * (size_t)L.kpStride * batch, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h->hDesc.p, h->dDesc.p, (size_t)L.kpStride * batch * 32, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int f = 0; f < batch; f++) {
        const int n = h->hCounts.p[f];
        if (n > kp_cap) return fail(h, ORBX_ERR_CAPACITY, "kp_cap too small");
        memcpy(kps + (size_t)f * kp_cap, h->hKps.p + (size_t)f * L.kpStride, sizeof(orbx_keypoint) * n);
        memcpy(desc + (size_t)f * kp_cap * 32, h->hDesc.p + (size_t)f * L.kpStride * 32, (size_t)n * 32);
        n_out[f] = n;
    }
    return ORBX_OK;
}

// OrbFrame::ComputeStereoMatches (orbframe.cpp:511-705) on the device-resident results of the last
// extraction of `left` (frame frame_left) and `right` (frame frame_right); the two may be one handle.
// OrbFrame::FilterKeyPoints on the device-resident results of the last extraction (frames frame0 .. frame0 + n_frames - 1)
int orbx_filter_keypoints(orbx_extractor *h, int frame0, int n_frames, const float box[4])
{
    if (!h) return ORBX_ERR_ARG;
    if (!box || n_frames < 1 || frame0 < 0 || frame0 + n_frames > h->lastBatch) return fail(h, ORBX_ERR_ARG, "bad argument or no previous extraction");
    if (!(box[1] > 2.0f)) return ORBX_OK;                 // orbframe.cpp:405: no bounding box configured, nothing is filtered
    CK(cudaSetDevice(h->cfg.device));
    CK(orderAfterLast(h, h->stream));
    CK(launch_filter_keypoints(h->dKps.p, h->dDesc.p, h->dCounts.p, h->L.kpStride, frame0, n_frames, box, h->stream));
    // the next extraction may run on another lane of this handle: it must not overwrite the records while they are compacted
    CK(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

int orbx_stereo_match(orbx_extractor *left, int frame_left, orbx_extractor *right, int frame_right, float mbf, float mb,
                      float *u_right, float *depth, int cap, int *n_left, int *n_matches)
{
    orbx_extractor *h = left;
    if (!left || !right) return ORBX_ERR_ARG;
    if (!u_right || !depth || frame_left < 0 || frame_left >= left->lastBatch || frame_right < 0 || frame_right >= right->lastBatch)
        return fail(h, ORBX_ERR_ARG, "bad argument or no previous extraction");
    if (left->cfg.device != right->cfg.device || left->curW != right->curW || left->curH != right->curH ||
        left->cfg.nlevels != right->cfg.nlevels || left->cfg.scale_factor != right->cfg.scale_factor)
        return fail(h, ORBX_ERR_ARG, "left and right extractors must share device, image size and pyramid");
    CK(cudaSetDevice(h->cfg.device));
    const OrbxLayout &L = left->L;
    CK(h->dStereo.ensure((size_t)2 * L.kpStride));
    CK(h->dStereoI.ensure((size_t)L.kpStride + 1));
    CK(h->hStereo.ensure((size_t)2 * L.kpStride + 2));
    if (right != left) { CK(drainLast(right)); CK(cudaStreamSynchronize(right->stream)); }
    cudaStream_t st = left->stream;
    CK(orderAfterLast(left, st));
    const float maxD = mbf / mb;            // orbframe.cpp:545-547 (minZ = mb; +inf when mb is still 0)
    float *dU = h->dStereo.p, *dD = h->dStereo.p + L.kpStride;
    int *dSad = h->dStereoI.p, *dN = h->dStereoI.p + L.kpStride;
    CK(launch_stereo(L, left->dPyr.p + (size_t)frame_left * L.slab, right->dPyr.p + (size_t)frame_right * right->L.slab,
                     left->dKps.p + (size_t)frame_left * L.kpStride, left->dDesc.p + (size_t)frame_left * L.kpStride * 32,
                     left->dCounts.p + frame_left, right->dKps.p + (size_t)frame_right * right->L.kpStride,
                     right->dDesc.p + (size_t)frame_right * right->L.kpStride * 32, right->dCounts.p + frame_right,
                     1, 0, mbf, maxD, dU, dD, dSad, dN, st));
    int counts[2] = {0, 0};
    CK(cudaMemcpyAsync(h->hStereo.p, h->dStereo.p, sizeof(float) * 2 * L.kpStride, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&counts[0], left->dCounts.p + frame_left, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&counts[1], dN, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (counts[0] > cap) return fail(h, ORBX_ERR_CAPACITY, "u_right / depth arrays too small");
    memcpy(u_right, h->hStereo.p, sizeof(float) * counts[0]);
    memcpy(depth, h->hStereo.p + L.kpStride, sizeof(float) * counts[0]);
    if (n_left) *n_left = counts[0];
    if (n_matches) *n_matches = counts[1];
    return ORBX_OK;
}

// ComputeStereoMatches for n_pairs pairs of ONE batch: pair p = frames (frame_left0 + p*frame_step, frame_right0 + p*frame_step)
int orbx_stereo_match_batch(orbx_extractor *h, int n_pairs, int frame_left0, int frame_right0, int frame_step, float mbf, float mb,
                            float *u_right, float *depth, int cap, int *n_left, int *n_matches)
{
    if (!h) return ORBX_ERR_ARG;
    if (!u_right || !depth || n_pairs < 1 || frame_step < 1 || frame_left0 < 0 || frame_right0 < 0 ||
        frame_left0 + (n_pairs - 1) * frame_step >= h->lastBatch || frame_right0 + (n_pairs - 1) * frame_step >= h->lastBatch)
        return fail(h, ORBX_ERR_ARG, "bad argument or pairs outside the last batch");
    CK(cudaSetDevice(h->cfg.device));
    const OrbxLayout &L = h->L;
    if (cap < L.kpStride) return fail(h, ORBX_ERR_CAPACITY, "u_right / depth rows must hold orbx_max_keypoints entries");
    const size_t rows = (size_t)n_pairs * L.kpStride;
    CK(h->dStereo.ensure(2 * rows));
    CK(h->dStereoI.ensure(rows + n_pairs));
    CK(h->hStereo.ensure(2 * rows + 2));
    cudaStream_t st = h->stream;
    CK(orderAfterLast(h, st));
    const float maxD = mbf / mb;
    float *dU = h->dStereo.p, *dD = h->dStereo.p + rows;
    int *dSad = h->dStereoI.p, *dN = h->dStereoI.p + rows;
    CK(launch_stereo(L, h->dPyr.p + (size_t)frame_left0 * L.slab, h->dPyr.p + (size_t)frame_right0 * L.slab,
                     h->dKps.p + (size_t)frame_left0 * L.kpStride, h->dDesc.p + (size_t)frame_left0 * L.kpStride * 32, h->dCounts.p + frame_left0,
                     h->dKps.p + (size_t)frame_right0 * L.kpStride, h->dDesc.p + (size_t)frame_right0 * L.kpStride * 32, h->dCounts.p + frame_right0,
                     n_pairs, frame_step, mbf, maxD, dU, dD, dSad, dN, st));
    CK(h->hCounts.ensure((size_t)std::max(h->cfg.max_batch, 2 * n_pairs)));
    std::vector<int> nm((size_t)n_pairs), nl((size_t)h->lastBatch);
    CK(cudaMemcpyAsync(h->hStereo.p, h->dStereo.p, sizeof(float) * 2 * rows, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(nl.data(), h->dCounts.p, sizeof(int) * (size_t)h->lastBatch, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(nm.data(), dN, sizeof(int) * (size_t)n_pairs, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int p = 0; p < n_pairs; p++) {
        const int n = nl[(size_t)frame_left0 + (size_t)p * frame_step];
        memcpy(u_right + (size_t)p * cap, h->hStereo.p + (size_t)p * L.kpStride, sizeof(float) * n);
        memcpy(depth + (size_t)p * cap, h->hStereo.p + rows + (size_t)p * L.kpStride, sizeof(float) * n);
        if (n_left) n_left[p] = n;
        if (n_matches) n_matches[p] = nm[p];
    }
    return ORBX_OK;
}

int orbx_extract(orbx_extractor *h, const uint8_t *img, int width, int height, size_t pitch,
                 orbx_keypoint *kps, int kp_cap, uint8_t *desc, int *n_out)
{
    const uint8_t *imgs[1] = {img};
    return orbx_extract_batch(h, imgs, 1, width, height, pitch, kps, kp_cap, desc, n_out);
}

int orbx_host_alloc(size_t bytes, void **ptr)
{
    if (!ptr || bytes == 0) return ORBX_ERR_ARG;
    *ptr = nullptr;
    const cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) { cudaGetLastError(); *ptr = nullptr; return e == cudaErrorMemoryAllocation ? ORBX_ERR_NOMEM : ORBX_ERR_CUDA; }
    return ORBX_OK;
}

void orbx_host_free(void *ptr) { if (ptr) cudaFreeHost(ptr); }

int orbx_get_level(orbx_extractor *h, int frame, int level, const uint8_t **host_ptr, int *width, int *height, size_t *pitch)
{
    if (!h) return ORBX_ERR_ARG;
    if (!host_ptr || frame < 0 || frame >= h->lastBatch || level < 0 || level >= h->L.nlevels)
        return fail(h, ORBX_ERR_ARG, "bad frame/level");
    CK(cudaSetDevice(h->cfg.device));
    const OrbxLevel &l = h->L.lv[level];
    // one pinned buffer holding all levels of all frames of the last call, filled lazily per level
    CK(h->hLevel.ensure((size_t)h->L.slab * h->cfg.max_batch));
    uint8_t *dst = h->hLevel.p + (size_t)frame * h->L.slab + l.off;
    CK(orderAfterLast(h, h->stream));
    CK(cudaMemcpyAsync(dst, h->dPyr.p + (size_t)frame * h->L.slab + l.off, (size_t)l.pitch * l.h, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *host_ptr = dst;
    if (width) *width = l.w;
    if (height) *height = l.h;
    if (pitch) *pitch = (size_t)l.pitch;
    return ORBX_OK;
}

int orbx_debug_blurred(orbx_extractor *h, int frame, int level, uint8_t *dst, size_t dst_bytes, int *width, int *height)
{
    if (!h) return ORBX_ERR_ARG;
    if (!dst || frame < 0 || frame >= h->lastBatch || level < 0 || level >= h->L.nlevels)
        return fail(h, ORBX_ERR_ARG, "bad frame/level");
    const OrbxLevel &l = h->L.lv[level];
    if (dst_bytes < (size_t)l.w * l.h) return fail(h, ORBX_ERR_CAPACITY, "dst too small");
    CK(cudaSetDevice(h->cfg.device));
    CK(drainLast(h));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy2D(dst, l.w, h->dBlur.p + (size_t)frame * h->L.slab + l.off, l.pitch, l.w, l.h, cudaMemcpyDeviceToHost));
    if (width) *width = l.w;
    if (height) *height = l.h;
    return ORBX_OK;
}

// Per-stage device times of the pipeline, serialised on the handle's stream with CUDA events:
// ms[0] resize chain, ms[1] FAST (incl. clearing the row summaries), ms[2] octree, ms[3] blur,
// ms[4] describe.  Re-runs the stages on the frames of the last call (level 0 is still resident).
int orbx_profile_stages(orbx_extractor *h, int reps, float *ms, int n_ms)
{
    if (!h) return ORBX_ERR_ARG;
    if (!ms || n_ms < 5 || reps < 1 || h->lastBatch < 1) return fail(h, ORBX_ERR_ARG, "bad argument or no previous call");
    CK(cudaSetDevice(h->cfg.device));
    const OrbxLayout &L = h->L;
    const int batch = h->lastBatch;
    cudaStream_t st = h->stream;
    CK(drainLast(h));
    cudaEvent_t ev[6];
    for (int i = 0; i < 6; i++) CK(cudaEventCreate(&ev[i]));
    for (int i = 0; i < 5; i++) ms[i] = 0.f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(ev[0], st));
        for (int l = 1; l < L.nlevels; l++) launch_resize(h->tmapsResize, 0, h->dPyr.p, L, l, (const int4 *)h->dRtab.p, batch, h->nSM, st);
        CK(cudaEventRecord(ev[1], st));
        CK(cudaMemsetAsync(h->dCnt.p, 0, (size_t)L.rowsPerFrame * batch * sizeof(uint32_t), st));
        CK(cudaMemsetAsync(h->dBest.p, 0, (size_t)L.rowsPerFrame * batch * sizeof(unsigned long long), st));
        CK(launch_fast(h->tmapsFast, 0, L, h->dSegs.p, 0, L.totalSegs, h->dCnt.p, h->dBest.p, nullptr, nullptr, 0, h->fastWinRows, h->fastListCap, batch, st));
        CK(cudaEventRecord(ev[2], st));
        CK(launch_octree(L, h->dCnt.p, h->dBest.p, h->dSlots.p, h->dLvlCount.p, h->maxRows, h->maxNodes, batch, st));
        CK(cudaEventRecord(ev[3], st));
        launch_blur(h->tmaps, h->dBlur.p, L, h->dTiles.p, (int)h->tiles.size(), h->taps, 0, batch, st);
        CK(cudaEventRecord(ev[4], st));
        launch_describe(h->tmapsDescA, h->tmapsDescB, 0, L, h->dSlots.p, h->dLvlCount.p, h->umax, h->dKps.p, h->dDesc.p, h->dCounts.p, batch, st);
        CK(cudaEventRecord(ev[5], st));
        CK(cudaStreamSynchronize(st));
        for (int i = 0; i < 5; i++) { float t = 0; CK(cudaEventElapsedTime(&t, ev[i], ev[i + 1])); ms[i] += t / reps; }
    }
    for (int i = 0; i < 6; i++) cudaEventDestroy(ev[i]);
    return ORBX_OK;
}

int orbx_debug_enable_candidates(orbx_extractor *h, int enable)
{
    if (!h) return ORBX_ERR_ARG;
    h->dbgEnabled = enable ? 1 : 0;
    return ORBX_OK;
}

int orbx_debug_candidates(orbx_extractor *h, int frame, int level, int *xs, int *ys, int *score, int cap)
{
    if (!h) return ORBX_ERR_ARG;
    if (!h->dbgEnabled || !h->dDbg.p || frame < 0 || frame >= h->lastBatch || level < 0 || level >= h->L.nlevels)
        return fail(h, ORBX_ERR_ARG, "candidate recording not enabled or bad frame/level");
    CK(cudaSetDevice(h->cfg.device));
    CK(drainLast(h));
    CK(cudaStreamSynchronize(h->stream));
    const int slot = frame * h->L.nlevels + level;
    int n = 0;
    CK(cudaMemcpy(&n, h->dDbgCount.p + slot, sizeof(int), cudaMemcpyDeviceToHost));
    if (n > h->dbgCap) return fail(h, ORBX_ERR_CAPACITY, "candidate debug buffer overflow");
    std::vector<OrbxDbgCand> tmp((size_t)std::max(n, 1));
    if (n > 0) CK(cudaMemcpy(tmp.data(), h->dDbg.p + (size_t)slot * h->dbgCap, sizeof(OrbxDbgCand) * n, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n && i < cap; i++) {
        xs[i] = tmp[i].xy & 0xffff; ys[i] = tmp[i].xy >> 16; score[i] = tmp[i].score;
    }
    return n;
}

} // extern "C"
