/*
 * orb_oracle.c -- TEST INFRASTRUCTURE ONLY (see orb_oracle.h).
 *
 * Plain-C CPU restatement of the reference's ORB front end.  Every function
 * cites the reference file:line it follows (paths under /root/reference) or,
 * for arithmetic that lives in OpenCV (not vendored in the reference), the
 * SURVEY.md Appendix-A model that was validated bit-for-bit against cv2 4.13.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (no -march, no -ffast-math),
 * mirroring the reference's CMakeLists.txt:37-50 so float32 expressions are
 * evaluated exactly like the reference binary's (SSE2 scalar, no FMA).
 */
#define _GNU_SOURCE
#include "orb_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* OpenCV scalar helpers: cvRound = round-half-even (SSE cvtss2si),    */
/* cvFloor, cvCeil.                                                     */
/* ------------------------------------------------------------------ */
static inline int cv_round_f(float v) { return (int)lrintf(v); }
static inline int cv_round_d(double v) { return (int)lrint(v); }
static inline int cv_floor_d(double v) { int i = (int)v; return i - (i > v); }
static inline int cv_ceil_d(double v) { int i = (int)v; return i + (i < v); }

static const int k_pattern[1024] = {
#include "orb_pattern.inc"
};
const int *orbo_pattern(void) { return k_pattern; }

static const int k_default_taps[7] = {18, 34, 48, 56, 48, 34, 18}; /* cv2 4.13, SURVEY A.4 */

/* ------------------------------------------------------------------ */
/* Constructor tables: orbextractor.cpp:476-548                         */
/* ------------------------------------------------------------------ */
int orbo_params_init(orbo_params *p, int nfeatures, float scale_factor, int nlevels,
                     int ini_th, int min_th, const int *taps7)
{
    if (!p || nlevels < 1 || nlevels > ORBO_MAX_LEVELS) return -1;
    memset(p, 0, sizeof(*p));
    p->nfeatures = nfeatures;
    p->scale_factor = scale_factor;
    p->nlevels = nlevels;
    p->ini_th = ini_th;
    p->min_th = min_th;
    memcpy(p->taps, taps7 ? taps7 : k_default_taps, sizeof(p->taps));

    /* :492-508 float chains */
    p->sf[0] = 1.0f;
    p->sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        p->sf[i] = p->sf[i - 1] * scale_factor;
        p->sigma2[i] = p->sf[i] * p->sf[i];
    }
    for (int i = 0; i < nlevels; i++) {
        p->inv_sf[i] = 1.0f / p->sf[i];
        p->inv_sigma2[i] = 1.0f / p->sigma2[i];
    }
    /* :512-523 per-level quotas */
    float factor = 1.0f / scale_factor;
    float ndes = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int level = 0; level < nlevels - 1; level++) {
        p->quota[level] = cv_round_f(ndes);
        sum += p->quota[level];
        ndes *= factor;
    }
    p->quota[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;

    /* :532-547 umax of the circular patch */
    const int HP = 15;
    int v, v0;
    int vmax = cv_floor_d(HP * sqrtf(2.f) / 2 + 1);
    int vmin = cv_ceil_d(HP * sqrtf(2.f) / 2);
    const double hp2 = HP * HP;
    for (v = 0; v <= vmax; ++v) p->umax[v] = cv_round_d(sqrt(hp2 - v * v));
    for (v = HP, v0 = 0; v >= vmin; --v) {
        while (p->umax[v0] == p->umax[v0 + 1]) ++v0;
        p->umax[v] = v0;
        ++v0;
    }
    return 0;
}

/* orbextractor.cpp:658-659 */
void orbo_level_size(const orbo_params *p, int w0, int h0, int level, int *w, int *h)
{
    float scale = p->inv_sf[level];
    *w = cv_round_f((float)w0 * scale);
    *h = cv_round_f((float)h0 * scale);
}

/* ------------------------------------------------------------------ */
/* cv::resize(INTER_LINEAR, 8UC1): SURVEY A.1 (call site :666)          */
/* ------------------------------------------------------------------ */
static void resize_axis_tab(int ssize, int dsize, int *ofs, short *c0, short *c1)
{
    double inv_scale = (double)dsize / ssize;
    double scale = 1.0 / inv_scale;
    for (int d = 0; d < dsize; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = cv_floor_d(f);
        f -= s;
        if (s < 0) { f = 0; s = 0; }
        if (s >= ssize - 1) { f = 0; s = ssize - 1; }
        ofs[d] = s;
        c0[d] = (short)cv_round_f((1.f - f) * 2048.f);
        c1[d] = (short)cv_round_f(f * 2048.f);
    }
}

void orbo_resize_linear_u8(const uint8_t *src, int sw, int sh, size_t sstride,
                           uint8_t *dst, int dw, int dh, size_t dstride)
{
    int *xo = malloc(sizeof(int) * dw), *yo = malloc(sizeof(int) * dh);
    short *xa = malloc(sizeof(short) * dw), *xb = malloc(sizeof(short) * dw);
    short *ya = malloc(sizeof(short) * dh), *yb = malloc(sizeof(short) * dh);
    int *r0 = malloc(sizeof(int) * dw), *r1 = malloc(sizeof(int) * dw);
    resize_axis_tab(sw, dw, xo, xa, xb);
    resize_axis_tab(sh, dh, yo, ya, yb);
    for (int y = 0; y < dh; y++) {
        int sy0 = yo[y], sy1 = sy0 + 1 < sh ? sy0 + 1 : sh - 1;
        const uint8_t *s0 = src + (size_t)sy0 * sstride, *s1 = src + (size_t)sy1 * sstride;
        for (int x = 0; x < dw; x++) {
            int sx0 = xo[x], sx1 = sx0 + 1 < sw ? sx0 + 1 : sw - 1;
            r0[x] = s0[sx0] * xa[x] + s0[sx1] * xb[x];
            r1[x] = s1[sx0] * xa[x] + s1[sx1] * xb[x];
        }
        int b0 = ya[y], b1 = yb[y];
        uint8_t *d = dst + (size_t)y * dstride;
        for (int x = 0; x < dw; x++)
            d[x] = (uint8_t)((((b0 * (r0[x] >> 4)) >> 16) + ((b1 * (r1[x] >> 4)) >> 16) + 2) >> 2);
    }
    free(xo); free(yo); free(xa); free(xb); free(ya); free(yb); free(r0); free(r1);
}

/* ------------------------------------------------------------------ */
/* cv::copyMakeBorder(BORDER_REFLECT_101): SURVEY A.2 (:668-674)        */
/* ------------------------------------------------------------------ */
static inline int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    }
    return p;
}

void orbo_border_reflect101_u8(const uint8_t *src, int w, int h, size_t sstride,
                               uint8_t *dst, size_t dstride, int top, int bottom, int left, int right)
{
    int dw = w + left + right, dh = h + top + bottom;
    /* read rows into a temporary first: dst may alias src (ROI of the same buffer) */
    uint8_t *tmp = malloc((size_t)w * h);
    for (int y = 0; y < h; y++) memcpy(tmp + (size_t)y * w, src + (size_t)y * sstride, w);
    for (int y = 0; y < dh; y++) {
        int sy = reflect101(y - top, h);
        uint8_t *d = dst + (size_t)y * dstride;
        for (int x = 0; x < dw; x++) d[x] = tmp[(size_t)sy * w + reflect101(x - left, w)];
    }
    free(tmp);
}

/* ------------------------------------------------------------------ */
/* cv::FAST(TYPE_9_16, nonmax=true): SURVEY A.3 (call sites :950-956)   */
/* ------------------------------------------------------------------ */
static const int k_ring_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int k_ring_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

/* corner score: max over the 16 contiguous 9-arcs of min(+-(Ip - Iring)) - 1 */
static int fast9_score(const uint8_t *p, size_t stride)
{
    int d[25];
    int v = p[0];
    for (int k = 0; k < 16; k++) d[k] = v - p[(ptrdiff_t)k_ring_dy[k] * (ptrdiff_t)stride + k_ring_dx[k]];
    for (int k = 16; k < 25; k++) d[k] = d[k - 16];
    int best = -256;
    for (int k = 0; k < 16; k++) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; j++) {
            if (d[k + j] < mn) mn = d[k + j];
            if (d[k + j] > mx) mx = d[k + j];
        }
        if (mn > best) best = mn;   /* centre brighter than the whole arc */
        if (-mx > best) best = -mx; /* centre darker than the whole arc   */
    }
    return best - 1;
}

/* Quick reject (same necessary condition OpenCV's scalar path tests first): every 9-arc
 * contains ring pixel k or k+8 for each k, so a corner at threshold t needs, for k = 0 and 4,
 * one of the pair brighter than v+t (or, for all pairs, darker than v-t). */
static inline int fast9_maybe(const uint8_t *p, const ptrdiff_t ofs[16], int t)
{
    int v = p[0], hi = v + t, lo = v - t;
    int a0 = p[ofs[0]], a8 = p[ofs[8]];
    int br = (a0 > hi) | (a8 > hi), dk = (a0 < lo) | (a8 < lo);
    if (!(br | dk)) return 0;
    int a4 = p[ofs[4]], a12 = p[ofs[12]];
    br &= (a4 > hi) | (a12 > hi);
    dk &= (a4 < lo) | (a12 < lo);
    if (!(br | dk)) return 0;
    int a2 = p[ofs[2]], a10 = p[ofs[10]], a6 = p[ofs[6]], a14 = p[ofs[14]];
    br &= ((a2 > hi) | (a10 > hi)) & ((a6 > hi) | (a14 > hi));
    dk &= ((a2 < lo) | (a10 < lo)) & ((a6 < lo) | (a14 < lo));
    return br | dk;
}

void orbo_fast9_score_map(const uint8_t *img, int w, int h, size_t stride, int threshold, uint8_t *score)
{
    ptrdiff_t ofs[16];
    for (int k = 0; k < 16; k++) ofs[k] = (ptrdiff_t)k_ring_dy[k] * (ptrdiff_t)stride + k_ring_dx[k];
    memset(score, 0, (size_t)w * h);
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            const uint8_t *p = img + (size_t)y * stride + x;
            if (!fast9_maybe(p, ofs, threshold)) continue;
            int s = fast9_score(p, stride);
            if (s >= threshold) score[(size_t)y * w + x] = (uint8_t)s;
        }
}

int orbo_fast9_nms(const uint8_t *img, int w, int h, size_t stride, int threshold,
                   int *xs, int *ys, int *score, int cap)
{
    if (w < 7 || h < 7) return 0;
    uint8_t *sm = malloc((size_t)w * h);
    orbo_fast9_score_map(img, w, h, stride, threshold, sm);
    int n = 0;
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            int s = sm[(size_t)y * w + x];
            if (s == 0) continue; /* not a corner (a score-0 corner can never beat its neighbours) */
            const uint8_t *c = sm + (size_t)y * w + x;
            /* strict > all 8 neighbours; outside the tested interior the map is 0 */
            if (s > c[-1] && s > c[1] && s > c[-w - 1] && s > c[-w] && s > c[-w + 1] &&
                s > c[w - 1] && s > c[w] && s > c[w + 1]) {
                if (n >= cap) { free(sm); return -1; }
                xs[n] = x; ys[n] = y; score[n] = s; n++;
            }
        }
    free(sm);
    return n;
}

/* ------------------------------------------------------------------ */
/* cv::GaussianBlur(7x7, sigma 2, REFLECT_101) u8: SURVEY A.4 (:622)    */
/* ------------------------------------------------------------------ */
void orbo_gaussian7_u8(const uint8_t *src, int w, int h, size_t sstride,
                       uint8_t *dst, size_t dstride, const int taps[7])
{
    /* horizontal pass into 16-bit rows (max 255*sum(taps) fits: sum <= 257), on a row padded
     * with its REFLECT_101 border; vertical pass over reflected row pointers. */
    uint16_t *hbuf = malloc(sizeof(uint16_t) * (size_t)w * h);
    uint8_t *row = malloc((size_t)w + 6);
    const uint32_t t0 = taps[0], t1 = taps[1], t2 = taps[2], t3 = taps[3], t4 = taps[4], t5 = taps[5], t6 = taps[6];
    for (int y = 0; y < h; y++) {
        const uint8_t *s = src + (size_t)y * sstride;
        for (int k = 0; k < 3; k++) {
            row[k] = s[reflect101(k - 3, w)];
            row[w + 3 + k] = s[reflect101(w + k, w)];
        }
        memcpy(row + 3, s, w);
        uint16_t *o = hbuf + (size_t)y * w;
        for (int x = 0; x < w; x++)
            o[x] = (uint16_t)(t0 * row[x] + t1 * row[x + 1] + t2 * row[x + 2] + t3 * row[x + 3] +
                              t4 * row[x + 4] + t5 * row[x + 5] + t6 * row[x + 6]);
    }
    for (int y = 0; y < h; y++) {
        const uint16_t *r[7];
        for (int k = 0; k < 7; k++) r[k] = hbuf + (size_t)reflect101(y + k - 3, h) * w;
        uint8_t *d = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            uint32_t acc = t0 * r[0][x] + t1 * r[1][x] + t2 * r[2][x] + t3 * r[3][x] +
                           t4 * r[4][x] + t5 * r[5][x] + t6 * r[6][x];
            uint32_t v = (acc + 32768u) >> 16;
            d[x] = (uint8_t)(v > 255 ? 255 : v);
        }
    }
    free(hbuf); free(row);
}

/* ------------------------------------------------------------------ */
/* cv::fastAtan2 scalar: SURVEY A.6 (call site :162)                    */
/* ------------------------------------------------------------------ */
float orbo_fast_atan2(float y, float x)
{
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale;
    const float p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale;
    const float p7 = -0.04432655554792128f * scale;
    float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)2.2204460492503131e-16);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)2.2204460492503131e-16);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* ------------------------------------------------------------------ */
/* Gridded FAST with per-cell threshold fallback: orbextractor.cpp:906-970 */
/* ------------------------------------------------------------------ */
int orbo_grid_fast(const uint8_t *lvl, int w, int h, size_t stride, int ini_th, int min_th,
                   int *xs, int *ys, int *score, int cap)
{
    const float W = 30;
    const int minBX = 16, minBY = 16; /* EDGE_THRESHOLD-3, :914-915 */
    const int maxBX = w - 19 + 3, maxBY = h - 19 + 3;
    const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    if (nCols <= 0 || nRows <= 0) return -2; /* reference divides by zero here */
    const int wCell = (int)ceilf(width / nCols), hCell = (int)ceilf(height / nRows);
    int cellcap = (wCell + 6) * (hCell + 6);
    int *cx = malloc(sizeof(int) * cellcap), *cy = malloc(sizeof(int) * cellcap), *cs = malloc(sizeof(int) * cellcap);
    int n = 0;
    for (int i = 0; i < nRows; i++) {
        const float iniY = (float)(minBY + i * hCell);
        float maxY = iniY + hCell + 6;
        if (iniY >= maxBY - 3) continue;
        if (maxY > maxBY) maxY = (float)maxBY;
        for (int j = 0; j < nCols; j++) {
            const float iniX = (float)(minBX + j * wCell);
            float maxX = iniX + wCell + 6;
            if (iniX >= maxBX - 6) continue;
            if (maxX > maxBX) maxX = (float)maxBX;
            int x0 = (int)iniX, x1 = (int)maxX, y0 = (int)iniY, y1 = (int)maxY;
            const uint8_t *sub = lvl + (size_t)y0 * stride + x0;
            int m = orbo_fast9_nms(sub, x1 - x0, y1 - y0, stride, ini_th, cx, cy, cs, cellcap);
            if (m == 0) m = orbo_fast9_nms(sub, x1 - x0, y1 - y0, stride, min_th, cx, cy, cs, cellcap);
            for (int k = 0; k < m; k++) {
                if (n >= cap) { free(cx); free(cy); free(cs); return -1; }
                xs[n] = cx[k] + j * wCell; /* :963-964 */
                ys[n] = cy[k] + i * hCell;
                score[n] = cs[k];
                n++;
            }
        }
    }
    free(cx); free(cy); free(cs);
    return n;
}

/* ------------------------------------------------------------------ */
/* DistributeOctTree: orbextractor.cpp:680-904, DivideNode :72-128      */
/* Faithful list simulation; vKeys hold candidate indices.              */
/* ------------------------------------------------------------------ */
typedef struct onode {
    struct onode *prev, *next;
    int *keys; int nkeys;
    int ULx, ULy, URx, URy, BLx, BLy, BRx, BRy;
    int no_more;
    long seq; /* creation order: stands in for the heap address under a monotone allocator */
} onode;

typedef struct { onode *head, *tail; int size; long next_seq; } olist;

static onode *onode_new(int cap)
{
    onode *n = calloc(1, sizeof(onode));
    n->keys = malloc(sizeof(int) * (cap > 0 ? cap : 1));
    return n;
}
static void onode_free(onode *n) { free(n->keys); free(n); }
static void olist_push_front(olist *l, onode *n)
{
    n->seq = l->next_seq++;
    n->prev = NULL; n->next = l->head;
    if (l->head) l->head->prev = n; else l->tail = n;
    l->head = n; l->size++;
}
static void olist_push_back(olist *l, onode *n)
{
    n->seq = l->next_seq++;
    n->next = NULL; n->prev = l->tail;
    if (l->tail) l->tail->next = n; else l->head = n;
    l->tail = n; l->size++;
}
static onode *olist_erase(olist *l, onode *n) /* returns next */
{
    onode *nx = n->next;
    if (n->prev) n->prev->next = n->next; else l->head = n->next;
    if (n->next) n->next->prev = n->prev; else l->tail = n->prev;
    l->size--;
    onode_free(n);
    return nx;
}

/* DivideNode :72-128 -- integer-division-inside-ceil kept as written */
static void divide_node(const onode *p, const int *xs, const int *ys, onode *c[4])
{
    const int halfX = (int)ceil((double)((p->URx - p->ULx) / 2));
    const int halfY = (int)ceil((double)((p->BRy - p->ULy) / 2));
    for (int i = 0; i < 4; i++) c[i] = onode_new(p->nkeys);
    onode *n1 = c[0], *n2 = c[1], *n3 = c[2], *n4 = c[3];
    n1->ULx = p->ULx; n1->ULy = p->ULy;
    n1->URx = p->ULx + halfX; n1->URy = p->ULy;
    n1->BLx = p->ULx; n1->BLy = p->ULy + halfY;
    n1->BRx = p->ULx + halfX; n1->BRy = p->ULy + halfY;

    n2->ULx = n1->URx; n2->ULy = n1->URy;
    n2->URx = p->URx; n2->URy = p->URy;
    n2->BLx = n1->BRx; n2->BLy = n1->BRy;
    n2->BRx = p->URx; n2->BRy = p->ULy + halfY;

    n3->ULx = n1->BLx; n3->ULy = n1->BLy;
    n3->URx = n1->BRx; n3->URy = n1->BRy;
    n3->BLx = p->BLx; n3->BLy = p->BLy;
    n3->BRx = n1->BRx; n3->BRy = p->BLy;

    n4->ULx = n3->URx; n4->ULy = n3->URy;
    n4->URx = n2->BRx; n4->URy = n2->BRy;
    n4->BLx = n3->BRx; n4->BLy = n3->BRy;
    n4->BRx = p->BRx; n4->BRy = p->BRy;

    for (int i = 0; i < p->nkeys; i++) {
        int k = p->keys[i];
        float px = (float)xs[k], py = (float)ys[k];
        onode *t;
        if (px < n1->URx) t = (py < n1->BRy) ? n1 : n3;
        else if (py < n1->BRy) t = n2;
        else t = n4;
        t->keys[t->nkeys++] = k;
    }
    for (int i = 0; i < 4; i++) if (c[i]->nkeys == 1) c[i]->no_more = 1;
}

typedef struct { int size; long seq; onode *node; } size_ptr;
static int size_ptr_cmp(const void *a, const void *b, void *arg)
{
    const size_ptr *x = a, *y = b;
    const int oldest_first = *(const int *)arg;
    if (x->size != y->size) return x->size < y->size ? -1 : 1;
    long sx = oldest_first ? -x->seq : x->seq, sy = oldest_first ? -y->seq : y->seq;
    return sx < sy ? -1 : (sx > sy ? 1 : 0);
}

/* push non-empty children n1..n4 to the front; record those with >1 key (:769-806, :833-866) */
static void push_children(olist *l, onode *c[4], size_ptr *rec, int *nrec, int *n_to_expand)
{
    for (int i = 0; i < 4; i++) {
        if (c[i]->nkeys > 0) {
            olist_push_front(l, c[i]);
            if (c[i]->nkeys > 1) {
                if (n_to_expand) (*n_to_expand)++;
                rec[*nrec].size = c[i]->nkeys;
                rec[*nrec].seq = c[i]->seq;
                rec[*nrec].node = c[i];
                (*nrec)++;
            }
        } else {
            onode_free(c[i]);
        }
    }
}

int orbo_distribute(const int *xs, const int *ys, const int *score, int n,
                    int minX, int maxX, int minY, int maxY, int N, int tie_rule,
                    int *out_idx, int out_cap)
{
    if (maxY - minY <= 0) return -2;
    /* :684-686 integer division before round */
    const int nIni = (int)round((double)((maxX - minX) / (maxY - minY)));
    if (nIni <= 0) return -2; /* reference: division by zero (portrait input) */
    const int hX = (maxX - minX) / nIni;

    olist l = {0};
    onode **ini = malloc(sizeof(onode *) * nIni);
    for (int i = 0; i < nIni; i++) {
        onode *ni = onode_new(n);
        ni->ULx = hX * i; ni->ULy = 0;
        ni->URx = hX * i + 1; ni->URy = 0; /* :697 as written */
        ni->BLx = ni->ULx; ni->BLy = maxY - minY;
        ni->BRx = ni->URx; ni->BRy = maxY - minY;
        olist_push_back(&l, ni);
        ini[i] = ni;
    }
    for (int i = 0; i < n; i++) {
        int s = (int)((float)xs[i] / hX); /* :710 */
        if (s < 0 || s >= nIni) { /* reference indexes out of bounds here */
            while (l.head) olist_erase(&l, l.head);
            free(ini);
            return -3;
        }
        ini[s]->keys[ini[s]->nkeys++] = i;
    }
    free(ini);
    for (onode *it = l.head; it;) { /* :715-726 */
        if (it->nkeys == 1) { it->no_more = 1; it = it->next; }
        else if (it->nkeys == 0) it = olist_erase(&l, it);
        else it = it->next;
    }

    int finish = 0;
    int reccap = 4 * (n + nIni) + 16;
    size_ptr *rec = malloc(sizeof(size_ptr) * reccap), *prev = malloc(sizeof(size_ptr) * reccap);
    int nrec = 0;
    while (!finish) {
        int prevSize = l.size;
        int nToExpand = 0;
        nrec = 0;
        for (onode *it = l.head; it;) {
            if (it->no_more) { it = it->next; continue; }
            onode *c[4];
            divide_node(it, xs, ys, c);
            push_children(&l, c, rec, &nrec, &nToExpand);
            it = olist_erase(&l, it);
        }
        if (l.size >= N || l.size == prevSize) {
            finish = 1;
        } else if (l.size + nToExpand * 3 > N) {
            while (!finish) {
                prevSize = l.size;
                int nprev = nrec;
                memcpy(prev, rec, sizeof(size_ptr) * nrec);
                nrec = 0;
                qsort_r(prev, nprev, sizeof(size_ptr), size_ptr_cmp, &tie_rule); /* :825 (size, pointer) */
                for (int j = nprev - 1; j >= 0; j--) {
                    onode *c[4];
                    divide_node(prev[j].node, xs, ys, c);
                    push_children(&l, c, rec, &nrec, NULL);
                    olist_erase(&l, prev[j].node);
                    if (l.size >= N) break;
                }
                if (l.size >= N || l.size == prevSize) finish = 1;
            }
        }
    }
    free(rec); free(prev);

    /* :885-901 best response per node, first wins ties */
    int cnt = 0, rc = 0;
    for (onode *it = l.head; it; it = it->next) {
        int best = it->keys[0];
        for (int k = 1; k < it->nkeys; k++)
            if ((float)score[it->keys[k]] > (float)score[best]) best = it->keys[k];
        if (cnt >= out_cap) { rc = -1; break; }
        out_idx[cnt++] = best;
    }
    while (l.head) olist_erase(&l, l.head);
    return rc < 0 ? rc : cnt;
}

/* ------------------------------------------------------------------ */
/* IC_Angle: orbextractor.cpp:136-163                                   */
/* ------------------------------------------------------------------ */
float orbo_ic_angle(const uint8_t *lvl, size_t stride, int cx, int cy, const int umax[16])
{
    int m_01 = 0, m_10 = 0;
    const uint8_t *center = lvl + (size_t)cy * stride + cx;
    for (int u = -15; u <= 15; ++u) m_10 += u * center[u];
    int step = (int)stride;
    for (int v = 1; v <= 15; ++v) {
        int v_sum = 0;
        int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int val_plus = center[u + v * step], val_minus = center[u - v * step];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return orbo_fast_atan2((float)m_01, (float)m_10);
}

/* ------------------------------------------------------------------ */
/* computeOrbDescriptor: orbextractor.cpp:165-203                       */
/* ------------------------------------------------------------------ */
void orbo_rbrief(const uint8_t *blurred, size_t stride, int cx, int cy, float angle_deg, uint8_t desc[32])
{
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.0);
    float angle = (float)angle_deg * factorPI;
    float a = (float)cosf(angle), b = (float)sinf(angle);
    const uint8_t *center = blurred + (size_t)cy * stride + cx;
    const int step = (int)stride;
    const int *pat = k_pattern;
#define ORBO_GET(idx) \
    center[cv_round_f(pat[2 * (idx)] * b + pat[2 * (idx) + 1] * a) * step + \
           cv_round_f(pat[2 * (idx)] * a - pat[2 * (idx) + 1] * b)]
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int k = 0; k < 8; k++) {
            int t0 = ORBO_GET(2 * k), t1 = ORBO_GET(2 * k + 1);
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
#undef ORBO_GET
}

/* ------------------------------------------------------------------ */
/* Whole extractor: ExtractFeatures orbextractor.cpp:582-642            */
/* ------------------------------------------------------------------ */
struct orbo_extractor {
    orbo_params p;
    int tie_rule;
    int w0, h0;
    /* pyramid: bordered buffers, :654-678 */
    uint8_t *buf[ORBO_MAX_LEVELS];
    int lw[ORBO_MAX_LEVELS], lh[ORBO_MAX_LEVELS];
    size_t lstride[ORBO_MAX_LEVELS];
    uint8_t *blur[ORBO_MAX_LEVELS];
    int has_blur[ORBO_MAX_LEVELS];
    int *cx[ORBO_MAX_LEVELS], *cy[ORBO_MAX_LEVELS], *cs[ORBO_MAX_LEVELS];
    int ncand[ORBO_MAX_LEVELS];
};

orbo_extractor *orbo_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th,
                            const int *taps7)
{
    orbo_extractor *e = calloc(1, sizeof(*e));
    if (orbo_params_init(&e->p, nfeatures, scale_factor, nlevels, ini_th, min_th, taps7) != 0) {
        free(e);
        return NULL;
    }
    return e;
}

static void orbo_release_levels(orbo_extractor *e)
{
    for (int l = 0; l < ORBO_MAX_LEVELS; l++) {
        free(e->buf[l]); e->buf[l] = NULL;
        free(e->blur[l]); e->blur[l] = NULL;
        free(e->cx[l]); free(e->cy[l]); free(e->cs[l]);
        e->cx[l] = e->cy[l] = e->cs[l] = NULL;
        e->ncand[l] = 0; e->has_blur[l] = 0;
    }
}

void orbo_destroy(orbo_extractor *e)
{
    if (!e) return;
    orbo_release_levels(e);
    free(e);
}

const orbo_params *orbo_get_params(const orbo_extractor *e) { return &e->p; }
void orbo_set_tie_rule(orbo_extractor *e, int tie_rule) { e->tie_rule = tie_rule; }

const uint8_t *orbo_level(const orbo_extractor *e, int level, int *w, int *h, size_t *stride)
{
    if (level < 0 || level >= e->p.nlevels || !e->buf[level]) return NULL;
    *w = e->lw[level]; *h = e->lh[level]; *stride = e->lstride[level];
    return e->buf[level] + 19 * e->lstride[level] + 19;
}

const uint8_t *orbo_blurred(const orbo_extractor *e, int level, int *w, int *h, size_t *stride)
{
    if (level < 0 || level >= e->p.nlevels || !e->has_blur[level]) return NULL;
    *w = e->lw[level]; *h = e->lh[level]; *stride = (size_t)e->lw[level];
    return e->blur[level];
}

int orbo_candidates(const orbo_extractor *e, int level, const int **xs, const int **ys, const int **score)
{
    if (level < 0 || level >= e->p.nlevels) return -1;
    *xs = e->cx[level]; *ys = e->cy[level]; *score = e->cs[level];
    return e->ncand[level];
}

int orbo_extract(orbo_extractor *e, const uint8_t *img, int w, int h, size_t stride,
                 orbo_keypoint *kps, uint8_t *desc, int cap)
{
    const orbo_params *p = &e->p;
    const int EDGE = 19;
    orbo_release_levels(e);
    e->w0 = w; e->h0 = h;

    /* ComputePyramid :654-678 */
    for (int l = 0; l < p->nlevels; l++) {
        int lw, lh;
        orbo_level_size(p, w, h, l, &lw, &lh);
        if (lw < 1 || lh < 1) return -2;
        e->lw[l] = lw; e->lh[l] = lh;
        e->lstride[l] = (size_t)lw + 2 * EDGE;
        e->buf[l] = malloc(e->lstride[l] * (size_t)(lh + 2 * EDGE));
        uint8_t *roi = e->buf[l] + EDGE * e->lstride[l] + EDGE;
        if (l != 0) {
            const uint8_t *prev = e->buf[l - 1] + EDGE * e->lstride[l - 1] + EDGE;
            orbo_resize_linear_u8(prev, e->lw[l - 1], e->lh[l - 1], e->lstride[l - 1], roi, lw, lh, e->lstride[l]);
            orbo_border_reflect101_u8(roi, lw, lh, e->lstride[l], e->buf[l], e->lstride[l], EDGE, EDGE, EDGE, EDGE);
        } else {
            orbo_border_reflect101_u8(img, w, h, stride, e->buf[l], e->lstride[l], EDGE, EDGE, EDGE, EDGE);
        }
    }

    /* ComputeKeyPointsOctTree :906-994 */
    int total = 0;
    int lvl_first[ORBO_MAX_LEVELS], lvl_count[ORBO_MAX_LEVELS];
    for (int l = 0; l < p->nlevels; l++) {
        int lw = e->lw[l], lh = e->lh[l];
        const uint8_t *roi = e->buf[l] + EDGE * e->lstride[l] + EDGE;
        int ccap = (lw > 38 && lh > 38) ? (lw - 38) * (lh - 38) : 1;
        e->cx[l] = malloc(sizeof(int) * ccap);
        e->cy[l] = malloc(sizeof(int) * ccap);
        e->cs[l] = malloc(sizeof(int) * ccap);
        int nc = orbo_grid_fast(roi, lw, lh, e->lstride[l], p->ini_th, p->min_th, e->cx[l], e->cy[l], e->cs[l], ccap);
        if (nc < 0) return nc;
        e->ncand[l] = nc;
        int *sel = malloc(sizeof(int) * (nc > 0 ? nc : 1));
        int ns = orbo_distribute(e->cx[l], e->cy[l], e->cs[l], nc, 16, lw - 16, 16, lh - 16,
                                 p->quota[l], e->tie_rule, sel, nc > 0 ? nc : 1);
        if (ns < 0) { free(sel); return ns; }
        if (total + ns > cap) { free(sel); return -1; }
        const int scaledPatchSize = 31 * (int)p->sf[l]; /* :978 int cast before multiply */
        lvl_first[l] = total; lvl_count[l] = ns;
        for (int i = 0; i < ns; i++) {
            orbo_keypoint *k = &kps[total + i];
            k->x = (float)e->cx[l][sel[i]] + 16; /* :984-985 */
            k->y = (float)e->cy[l][sel[i]] + 16;
            k->size = (float)scaledPatchSize;
            k->response = (float)e->cs[l][sel[i]];
            k->octave = l;
            k->class_id = -1;
            k->angle = -1;
        }
        total += ns;
        free(sel);
    }
    /* computeOrientation :992-993 */
    for (int l = 0; l < p->nlevels; l++) {
        const uint8_t *roi = e->buf[l] + EDGE * e->lstride[l] + EDGE;
        for (int i = 0; i < lvl_count[l]; i++) {
            orbo_keypoint *k = &kps[lvl_first[l] + i];
            k->angle = orbo_ic_angle(roi, e->lstride[l], cv_round_f(k->x), cv_round_f(k->y), p->umax);
        }
    }
    /* descriptors :612-640 */
    for (int l = 0; l < p->nlevels; l++) {
        if (lvl_count[l] == 0) continue;
        int lw = e->lw[l], lh = e->lh[l];
        const uint8_t *roi = e->buf[l] + EDGE * e->lstride[l] + EDGE;
        uint8_t *work = malloc((size_t)lw * lh); /* clone(): contiguous copy of the ROI */
        for (int y = 0; y < lh; y++) memcpy(work + (size_t)y * lw, roi + (size_t)y * e->lstride[l], lw);
        e->blur[l] = malloc((size_t)lw * lh);
        orbo_gaussian7_u8(work, lw, lh, lw, e->blur[l], lw, p->taps);
        free(work);
        e->has_blur[l] = 1;
        for (int i = 0; i < lvl_count[l]; i++) {
            orbo_keypoint *k = &kps[lvl_first[l] + i];
            orbo_rbrief(e->blur[l], lw, cv_round_f(k->x), cv_round_f(k->y), k->angle, desc + (size_t)(lvl_first[l] + i) * 32);
        }
        if (l != 0) {
            float scale = p->sf[l];
            for (int i = 0; i < lvl_count[l]; i++) {
                orbo_keypoint *k = &kps[lvl_first[l] + i];
                k->x *= scale; k->y *= scale;
            }
        }
    }
    return total;
}

/* frame-partitioned multi-threaded driver for the CPU baseline (one extractor per thread,
 * mirrors the reference running independent instances in separate threads, orbframe.cpp:73-76) */
typedef struct {
    const orbo_params *p; int tie_rule;
    const uint8_t *const *imgs; int nimg, w, h; size_t stride;
    orbo_keypoint *kps; uint8_t *desc; int cap; int *counts;
    int tid, nthreads;
} extract_job;

static void *extract_worker(void *arg)
{
    extract_job *j = arg;
    orbo_extractor *e = orbo_create(j->p->nfeatures, j->p->scale_factor, j->p->nlevels, j->p->ini_th, j->p->min_th, j->p->taps);
    e->tie_rule = j->tie_rule;
    for (int i = j->tid; i < j->nimg; i += j->nthreads)
        j->counts[i] = orbo_extract(e, j->imgs[i], j->w, j->h, j->stride,
                                    j->kps + (size_t)i * j->cap, j->desc + (size_t)i * j->cap * 32, j->cap);
    orbo_destroy(e);
    return NULL;
}

int orbo_extract_batch_mt(const orbo_extractor *cfg, const uint8_t *const *imgs, int nimg, int w, int h, size_t stride,
                          orbo_keypoint *kps, uint8_t *desc, int cap, int *counts, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
    extract_job *jobs = malloc(sizeof(extract_job) * nthreads);
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (extract_job){&cfg->p, cfg->tie_rule, imgs, nimg, w, h, stride, kps, desc, cap, counts, t, nthreads};
        pthread_create(&th[t], NULL, extract_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
    return 0;
}

/* ------------------------------------------------------------------ */
/* OrbFrame::ComputeStereoMatches, orbframe.cpp:511-705 ("next" row N1). */
/* PINNED: the reference's own src/orbframe.cpp (stereo constructor +     */
/* ComputeStereoMatches) compiles unmodified against oracle/cvshim        */
/* (oracle/_ref/libframeref.so); mvuRight / m_depths equal this            */
/* restatement bit for bit (tests/test_oracle_vs_ref.py,                   */
/* tests/golden/ref_stereo.npz).                                           */
/* ------------------------------------------------------------------ */
typedef struct { int dist, idx; } dist_idx;
static int dist_idx_cmp(const void *a, const void *b)
{
    const dist_idx *x = a, *y = b;
    if (x->dist != y->dist) return x->dist < y->dist ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}

int orbo_stereo_matches(const orbo_keypoint *kl, const uint8_t *dl, int nl,
                        const orbo_keypoint *kr, const uint8_t *dr, int nr,
                        const uint8_t *const *pyrL, const uint8_t *const *pyrR, const int *lw, const int *lh,
                        const size_t *lstride, const float *sf, const float *inv_sf,
                        float mbf, float mb, float *uRight, float *depth)
{
    const int TH_HIGH = 100, TH_LOW = 50;               /* orbmatcher.cpp:36-37 */
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;       /* :516 */
    const int nRows = lh[0];                            /* :518 */
    for (int i = 0; i < nl; i++) { uRight[i] = -1.0f; depth[i] = -1.0f; }
    /* row table :521-543 */
    int *rowCount = calloc(nRows + 1, sizeof(int));
    for (int iR = 0; iR < nr; iR++) {
        const float kpY = kr[iR].y;
        const float r = 2.0f * sf[kr[iR].octave];
        const int maxr = (int)ceilf(kpY + r), minr = (int)floorf(kpY - r);
        for (int yi = minr; yi <= maxr; yi++) if (yi >= 0 && yi < nRows) rowCount[yi]++;
    }
    int *rowStart = malloc(sizeof(int) * (nRows + 1));
    rowStart[0] = 0;
    for (int i = 0; i < nRows; i++) rowStart[i + 1] = rowStart[i] + rowCount[i];
    int *rowIdx = malloc(sizeof(int) * (rowStart[nRows] + 1));
    memset(rowCount, 0, sizeof(int) * (nRows + 1));
    for (int iR = 0; iR < nr; iR++) {
        const float kpY = kr[iR].y;
        const float r = 2.0f * sf[kr[iR].octave];
        const int maxr = (int)ceilf(kpY + r), minr = (int)floorf(kpY - r);
        for (int yi = minr; yi <= maxr; yi++) if (yi >= 0 && yi < nRows) rowIdx[rowStart[yi] + rowCount[yi]++] = iR;
    }
    const float minZ = mb, minD = 0, maxD = mbf / minZ;  /* :545-547 */
    dist_idx *vDistIdx = malloc(sizeof(dist_idx) * (nl > 0 ? nl : 1));
    int nDist = 0;
    for (int iL = 0; iL < nl; iL++) {
        const orbo_keypoint *kpL = &kl[iL];
        const int levelL = kpL->octave;
        const float vL = kpL->y, uL = kpL->x;
        const int row = (int)vL;
        if (row < 0 || row >= nRows || rowCount[row] == 0) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH;
        int bestIdxR = 0;
        for (int iC = 0; iC < rowCount[row]; iC++) {
            const int iR = rowIdx[rowStart[row] + iC];
            const orbo_keypoint *kpR = &kr[iR];
            if (kpR->octave < levelL - 1 || kpR->octave > levelL + 1) continue;
            const float uR = kpR->x;
            if (uR >= minU && uR <= maxU) {
                const int dist = orbo_descriptor_distance(dl + (size_t)iL * 32, dr + (size_t)iR * 32);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist < thOrbDist) {                       /* :603 sub-pixel match by correlation */
            const float uR0 = kr[bestIdxR].x;
            const float scaleFactor = inv_sf[kpL->octave];
            const float scaleduL = roundf(kpL->x * scaleFactor);
            const float scaledvL = roundf(kpL->y * scaleFactor);
            const float scaleduR0 = roundf(uR0 * scaleFactor);
            const int w = 5, L = 5;
            const int lv = kpL->octave;
            const uint8_t *IL = pyrL[lv] + (size_t)((int)(scaledvL - w)) * lstride[lv] + (int)(scaleduL - w);
            const float cL = (float)IL[(size_t)w * lstride[lv] + w];
            int bestSad = 0x7fffffff;
            int bestincR = 0;
            float vDists[11];
            const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
            if (iniu < 0 || endu >= lw[lv]) continue;
            for (int incR = -L; incR <= +L; incR++) {
                const uint8_t *IR = pyrR[lv] + (size_t)((int)(scaledvL - w)) * lstride[lv] + (int)(scaleduR0 + incR - w);
                const float cR = (float)IR[(size_t)w * lstride[lv] + w];
                double acc = 0;                            /* cv::norm(IL, IR, NORM_L1) on CV_32F */
                for (int y = 0; y < 2 * w + 1; y++)
                    for (int x = 0; x < 2 * w + 1; x++) {
                        const float a = (float)IL[(size_t)y * lstride[lv] + x] - cL;
                        const float b = (float)IR[(size_t)y * lstride[lv] + x] - cR;
                        acc += fabs((double)(a - b));
                    }
                const float dist = (float)acc;
                if (dist < (float)bestSad) { bestSad = (int)dist; bestincR = incR; }
                vDists[L + incR] = dist;
            }
            if (bestincR == -L || bestincR == L) continue;
            const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = sf[kpL->octave] * (scaleduR0 + (float)bestincR + deltaR);
            float disparity = (uL - bestuR);
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) { disparity = 0.01f; bestuR = (float)(uL - 0.01); }
                depth[iL] = mbf / disparity;
                uRight[iL] = bestuR;
                vDistIdx[nDist].dist = bestSad; vDistIdx[nDist].idx = iL; nDist++;
            }
        }
    }
    if (nDist > 0) {                                      /* :693-705 (the reference reads vDistIdx[0] of an empty vector) */
        qsort(vDistIdx, nDist, sizeof(dist_idx), dist_idx_cmp);
        const float median = (float)vDistIdx[nDist / 2].dist;
        const float thDist = 1.5f * 1.4f * median;
        for (int i = nDist - 1; i >= 0; i--) {
            if ((float)vDistIdx[i].dist < thDist) break;
            uRight[vDistIdx[i].idx] = -1; depth[vDistIdx[i].idx] = -1;
        }
    }
    free(rowCount); free(rowStart); free(rowIdx); free(vDistIdx);
    return nDist;
}

/* ------------------------------------------------------------------ */
/* Matcher: orbmatcher.cpp:1662-1677 and the best/second loop :208-232  */
/* ------------------------------------------------------------------ */
int orbo_descriptor_distance(const uint8_t a[32], const uint8_t b[32])
{
    int32_t pa[8], pb[8];
    memcpy(pa, a, 32); memcpy(pb, b, 32);
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        unsigned int v = (unsigned int)(pa[i] ^ pb[i]);
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

void orbo_knn2(const uint8_t *q, int nq, const uint8_t *t, int nt,
               int32_t *idx, int32_t *d1, int32_t *d2)
{
    for (int i = 0; i < nq; i++) {
        int best1 = 256, best2 = 256, bi = -1; /* :208-210 */
        for (int j = 0; j < nt; j++) {
            int dist = orbo_descriptor_distance(q + (size_t)i * 32, t + (size_t)j * 32);
            if (dist < best1) { best2 = best1; best1 = dist; bi = j; }
            else if (dist < best2) { best2 = dist; }
        }
        idx[i] = bi; d1[i] = best1; d2[i] = best2;
    }
}

/* ------------------------------------------------------------------ */
/* ORBmatcher::SearchByProjection(frame, map points, th), orbmatcher.cpp:42-124, with the frame side it calls:       */
/* OrbFrame::AssignFeaturesToGrid / PosInGrid (orbframe.cpp:192-211, :381-393) and GetFeaturesInArea (:308-380).    */
/* PINNED: the reference's own translation units run the same search in oracle/_ref/libframeref.so                  */
/* (tests/test_oracle_vs_ref.py, tests/golden/ref_projection.npz).  Floats are combined one operation at a time in  */
/* the reference's order (this file is built with -ffp-contract=off).                                               */
/* ------------------------------------------------------------------ */
#define OG_COLS 64
#define OG_ROWS 48
#define OMAX(a, b) ((a) > (b) ? (a) : (b))
#define OMIN(a, b) ((a) < (b) ? (a) : (b))
int orbo_search_by_projection(const orbo_keypoint *keys, const float *uright, const uint8_t *occupied, const uint8_t *desc, int n,
                              float min_x, float min_y, float max_x, float max_y,
                              const uint8_t *mp_desc, const float *mp_x, const float *mp_y, const int32_t *mp_level,
                              const float *mp_radius, const uint8_t *mp_observed, int n_mp, float nnratio, int th_high,
                              int32_t *mp_match, int32_t *assigned)
{
    const float invW = (float)OG_COLS / (max_x - min_x), invH = (float)OG_ROWS / (max_y - min_y);   /* orbframe.cpp:179-180 */
    /* live copy of "F->m_mapPoints[idx] is set and has observations" (orbmatcher.cpp:87-89): the loop is sequential, an accepted
     * map point is stored at once (:121) and, when it has observations (mp_observed[i]), hides its key point from later ones */
    uint8_t *occ = calloc(n > 0 ? n : 1, 1);
    if (occupied) memcpy(occ, occupied, (size_t)n);
    /* AssignFeaturesToGrid: m_grid[ix][iy] lists in key-point order, :202-209 */
    int *count = calloc(OG_COLS * OG_ROWS + 1, sizeof(int)), *cell = malloc(sizeof(int) * (n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
        const int px = (int)round((keys[i].x - min_x) * invW), py = (int)round((keys[i].y - min_y) * invH);   /* :383-384 */
        cell[i] = (px < 0 || px >= OG_COLS || py < 0 || py >= OG_ROWS) ? -1 : px * OG_ROWS + py;
        if (cell[i] >= 0) count[cell[i] + 1]++;
    }
    for (int c = 0; c < OG_COLS * OG_ROWS; c++) count[c + 1] += count[c];
    int *fill = malloc(sizeof(int) * OG_COLS * OG_ROWS), *items = malloc(sizeof(int) * (n > 0 ? n : 1));
    memcpy(fill, count, sizeof(int) * OG_COLS * OG_ROWS);
    for (int i = 0; i < n; i++) if (cell[i] >= 0) items[fill[cell[i]]++] = i;
    for (int k = 0; k < n; k++) assigned[k] = -1;
    int nmatches = 0;
    for (int i = 0; i < n_mp; i++) {
        const float x = mp_x[i], y = mp_y[i], r = mp_radius[i];
        const int minLevel = mp_level[i] - 1, maxLevel = mp_level[i];                  /* orbmatcher.cpp:67-68 */
        mp_match[i] = -1;
        const int c0x = OMAX(0, (int)floor((x - min_x - r) * invW));                  /* orbframe.cpp:313-335 */
        if (c0x >= OG_COLS) continue;
        const int c1x = OMIN(OG_COLS - 1, (int)ceil((x - min_x + r) * invW));
        if (c1x < 0) continue;
        const int c0y = OMAX(0, (int)floor((y - min_y - r) * invH));
        if (c0y >= OG_ROWS) continue;
        const int c1y = OMIN(OG_ROWS - 1, (int)ceil((y - min_y + r) * invH));
        if (c1y < 0) continue;
        const int checkLevels = (minLevel > 0) || (maxLevel >= 0);
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int ix = c0x; ix <= c1x; ix++)
            for (int iy = c0y; iy <= c1y; iy++)
                for (int e = count[ix * OG_ROWS + iy]; e < count[ix * OG_ROWS + iy + 1]; e++) {
                    const int idx = items[e];
                    if (checkLevels) {
                        if (keys[idx].octave < minLevel) continue;
                        if (maxLevel >= 0 && keys[idx].octave > maxLevel) continue;
                    }
                    const float distx = keys[idx].x - x, disty = keys[idx].y - y;
                    if (!(fabs(distx) < r && fabs(disty) < r)) continue;               /* :370 */
                    if (occ[idx]) continue;                                            /* orbmatcher.cpp:87-89 */
                    if (uright[idx] > 0) {
                        const float er = (float)fabs(x - uright[idx]);                 /* :93 */
                        if (er > r) continue;
                    }
                    const int dist = orbo_descriptor_distance(mp_desc + (size_t)i * 32, desc + (size_t)idx * 32);
                    if (dist < bestDist) {
                        bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = keys[idx].octave; bestIdx = idx;
                    } else if (dist < bestDist2) {
                        bestLevel2 = keys[idx].octave; bestDist2 = dist;
                    }
                }
        if (bestDist <= th_high) {                                                     /* :116-123 */
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            mp_match[i] = bestIdx;
            assigned[bestIdx] = i;                                                     /* :121 */
            occ[bestIdx] = (uint8_t)(mp_observed ? (mp_observed[i] != 0) : 0);         /* what :87-89 sees from now on */
            nmatches++;
        }
    }
    free(occ);
    free(count); free(cell); free(fill); free(items);
    return nmatches;
}

/* OrbFrame::FilterKeyPoints (orbframe.cpp:403-445): key points strictly inside box = {x0, x1, y0, y1} are dropped, the others
 * keep their order (in place); nothing happens unless box[1] > 2 (:405).  Returns the new count.
 * PINNED through the reference's own OrbFrame constructor (tests/test_oracle_vs_ref.py, tests/golden/ref_stereo.npz). */
int orbo_filter_keypoints(orbo_keypoint *keys, uint8_t *desc, int n, const float box[4])
{
    if (!(box[1] > 2)) return n;
    int v = 0;
    for (int i = 0; i < n; i++) {
        int inBounds = keys[i].x > box[0] && keys[i].x < box[1];
        inBounds &= keys[i].y > box[2] && keys[i].y < box[3];
        if (!inBounds) {
            keys[v] = keys[i];
            memmove(desc + (size_t)v * 32, desc + (size_t)i * 32, 32);
            v++;
        }
    }
    return v;
}

/* OrbFrame::AssignFeaturesToGrid (orbframe.cpp:192-211) + PosInGrid (:381-393) as CSR: cell ix * 48 + iy = m_grid[ix][iy]. */
void orbo_assign_grid(const orbo_keypoint *keys, int n, float min_x, float min_y, float max_x, float max_y,
                      int32_t *cell_start, int32_t *cell_items)
{
    const float invW = (float)OG_COLS / (max_x - min_x), invH = (float)OG_ROWS / (max_y - min_y);
    int *cell = malloc(sizeof(int) * (n > 0 ? n : 1));
    for (int c = 0; c <= OG_COLS * OG_ROWS; c++) cell_start[c] = 0;
    for (int i = 0; i < n; i++) {
        const int px = (int)round((keys[i].x - min_x) * invW), py = (int)round((keys[i].y - min_y) * invH);
        cell[i] = (px < 0 || px >= OG_COLS || py < 0 || py >= OG_ROWS) ? -1 : px * OG_ROWS + py;
        if (cell[i] >= 0) cell_start[cell[i] + 1]++;
    }
    for (int c = 0; c < OG_COLS * OG_ROWS; c++) cell_start[c + 1] += cell_start[c];
    int *fill = malloc(sizeof(int) * OG_COLS * OG_ROWS);
    memcpy(fill, cell_start, sizeof(int) * OG_COLS * OG_ROWS);
    for (int i = 0; i < n; i++) if (cell[i] >= 0) cell_items[fill[cell[i]]++] = i;
    free(cell); free(fill);
}

/* OrbFrame::GetFeaturesInArea (orbframe.cpp:308-380) for nq windows over the grid of AssignFeaturesToGrid (:192-211);
 * offsets[nq + 1] / indices in the reference's order, dist = DescriptorDistance to q_desc[i] when q_desc != NULL.
 * Returns the number of entries (nothing is written beyond cap). */
int orbo_area_distances(const orbo_keypoint *keys, const uint8_t *desc, int n, float min_x, float min_y, float max_x, float max_y,
                        const uint8_t *q_desc, const float *q_x, const float *q_y, const float *q_r, const int32_t *q_min_level,
                        const int32_t *q_max_level, int nq, int32_t *offsets, int32_t *indices, int32_t *dist, int cap)
{
    const float invW = (float)OG_COLS / (max_x - min_x), invH = (float)OG_ROWS / (max_y - min_y);
    int *count = calloc(OG_COLS * OG_ROWS + 1, sizeof(int)), *cell = malloc(sizeof(int) * (n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
        const int px = (int)round((keys[i].x - min_x) * invW), py = (int)round((keys[i].y - min_y) * invH);
        cell[i] = (px < 0 || px >= OG_COLS || py < 0 || py >= OG_ROWS) ? -1 : px * OG_ROWS + py;
        if (cell[i] >= 0) count[cell[i] + 1]++;
    }
    for (int c = 0; c < OG_COLS * OG_ROWS; c++) count[c + 1] += count[c];
    int *fill = malloc(sizeof(int) * OG_COLS * OG_ROWS), *items = malloc(sizeof(int) * (n > 0 ? n : 1));
    memcpy(fill, count, sizeof(int) * OG_COLS * OG_ROWS);
    for (int i = 0; i < n; i++) if (cell[i] >= 0) items[fill[cell[i]]++] = i;
    int total = 0;
    offsets[0] = 0;
    for (int i = 0; i < nq; i++) {
        const float x = q_x[i], y = q_y[i], r = q_r[i];
        const int minLevel = q_min_level[i], maxLevel = q_max_level[i];
        const int c0x = OMAX(0, (int)floor((x - min_x - r) * invW));
        const int c1x = OMIN(OG_COLS - 1, (int)ceil((x - min_x + r) * invW));
        const int c0y = OMAX(0, (int)floor((y - min_y - r) * invH));
        const int c1y = OMIN(OG_ROWS - 1, (int)ceil((y - min_y + r) * invH));
        if (!(c0x >= OG_COLS || c1x < 0 || c0y >= OG_ROWS || c1y < 0)) {               /* the four early returns, :314-335 */
            const int checkLevels = (minLevel > 0) || (maxLevel >= 0);
            for (int ix = c0x; ix <= c1x; ix++)
                for (int iy = c0y; iy <= c1y; iy++)
                    for (int e = count[ix * OG_ROWS + iy]; e < count[ix * OG_ROWS + iy + 1]; e++) {
                        const int idx = items[e];
                        if (checkLevels) {
                            if (keys[idx].octave < minLevel) continue;
                            if (maxLevel >= 0 && keys[idx].octave > maxLevel) continue;
                        }
                        const float distx = keys[idx].x - x, disty = keys[idx].y - y;
                        if (!(fabs(distx) < r && fabs(disty) < r)) continue;
                        if (total < cap) {
                            indices[total] = idx;
                            if (q_desc) dist[total] = orbo_descriptor_distance(q_desc + (size_t)i * 32, desc + (size_t)idx * 32);
                        }
                        total++;
                    }
        }
        offsets[i + 1] = total;
    }
    free(count); free(cell); free(fill); free(items);
    return total;
}

/* candidate-list variant: the inner loop of SearchByProjection, orbmatcher.cpp:76-114.  The reference
 * carries the octave of the best and of the second-best candidate (bestLevel, bestLevel2); the index of
 * the second best is returned instead so the caller can look the octave up. */
void orbo_knn2_csr(const uint8_t *q, int nq, const uint8_t *t, const int32_t *offsets, const int32_t *indices,
                   int32_t *idx1, int32_t *d1, int32_t *idx2, int32_t *d2)
{
    for (int i = 0; i < nq; i++) {
        int bestDist = 256, bestDist2 = 256, bestIdx = -1, bestIdx2 = -1;
        for (int k = offsets[i]; k < offsets[i + 1]; k++) {
            const int idx = indices[k];
            const int dist = orbo_descriptor_distance(q + (size_t)i * 32, t + (size_t)idx * 32);
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = bestIdx; bestIdx = idx; }
            else if (dist < bestDist2) { bestIdx2 = idx; bestDist2 = dist; }
        }
        idx1[i] = bestIdx; d1[i] = bestDist; idx2[i] = bestIdx2; d2[i] = bestDist2;
    }
}

/* OrbMapPoint::ComputeDistinctiveDescriptors, orbmappoint.cpp:314-383, for n_points map points at once: point p
 * observes the descriptor rows indices[offsets[p] .. offsets[p+1]) of `desc` (the reference gathers them from the
 * key frames, :328-337).  All-pairs DescriptorDistance (:350-358), per row the median = element
 * (int)(0.5*(N-1)) of the sorted row including the zero on the diagonal (:367), first row with the least median
 * wins (:369-373).  best[p] = position inside the point's list, -1 for an empty list (the reference returns early). */
static int cmp_int(const void *a, const void *b) { return *(const int *)a - *(const int *)b; }
void orbo_distinctive(const uint8_t *desc, const int32_t *offsets, const int32_t *indices, int n_points,
                      int32_t *best, int32_t *median)
{
    for (int p = 0; p < n_points; p++) {
        const int N = offsets[p + 1] - offsets[p];
        const int32_t *ix = indices + offsets[p];
        best[p] = -1; if (median) median[p] = -1;
        if (N <= 0) continue;
        float *dist = malloc(sizeof(float) * (size_t)N * N);
        int *row = malloc(sizeof(int) * (size_t)N);
        for (int i = 0; i < N; i++) {
            dist[(size_t)i * N + i] = 0;
            for (int j = i + 1; j < N; j++) {
                const int d = orbo_descriptor_distance(desc + (size_t)ix[i] * 32, desc + (size_t)ix[j] * 32);
                dist[(size_t)i * N + j] = (float)d; dist[(size_t)j * N + i] = (float)d;
            }
        }
        int bestMedian = 0x7fffffff, bestIdx = 0;
        for (int i = 0; i < N; i++) {
            for (int j = 0; j < N; j++) row[j] = (int)dist[(size_t)i * N + j];
            qsort(row, (size_t)N, sizeof(int), cmp_int);
            const int med = row[(int)(0.5 * ((float)N - 1.0))];
            if (med < bestMedian) { bestMedian = med; bestIdx = i; }
        }
        best[p] = bestIdx; if (median) median[p] = bestMedian;
        free(dist); free(row);
    }
}

/* OrbVocabulary::transform5, orbvocabulary.cpp:203-242, for n features.  The tree is given as arrays: node 0 is the
 * root, the children of node v are child_ids[child_off[v] .. child_off[v+1]) in the reference's order
 * (m_nodes[v].children), a node without children is a leaf (isLeaf) carrying word_id[v]; node_desc[v] is its
 * 32-byte descriptor.  Distances are OrbDescriptor::distance (orbdescriptor.cpp:75-95, the same popcount as
 * DescriptorDistance); the first child with the least distance is followed (strict '<', :224-232).
 * node[i] = the node passed at level L - levels_up (0 = root when that level is <= 0, :211-212). */
void orbo_voc_transform(const int32_t *child_off, const int32_t *child_ids, const uint8_t *node_desc, const int32_t *word_id,
                        int L, int levels_up, const uint8_t *feat, int n, int32_t *word, int32_t *node)
{
    const int nodeLevel = L - levels_up;
    for (int i = 0; i < n; i++) {
        const uint8_t *f = feat + (size_t)i * 32;
        int32_t nid = 0, fin = 0;
        int level = 0;
        do {
            ++level;
            const int32_t *ch = child_ids + child_off[fin];
            const int nc = child_off[fin + 1] - child_off[fin];
            fin = ch[0];
            double best = orbo_descriptor_distance(f, node_desc + (size_t)fin * 32);
            for (int c = 1; c < nc; c++) {
                const double d = orbo_descriptor_distance(f, node_desc + (size_t)ch[c] * 32);
                if (d < best) { best = d; fin = ch[c]; }
            }
            if (level == nodeLevel) nid = fin;
        } while (child_off[fin + 1] - child_off[fin] > 0);
        word[i] = word_id[fin];
        node[i] = nid;
    }
}

typedef struct { const uint8_t *q, *t; int q0, q1, nt; int32_t *idx, *d1, *d2; } knn_job;
static void *knn_worker(void *arg)
{
    knn_job *j = arg;
    orbo_knn2(j->q + (size_t)j->q0 * 32, j->q1 - j->q0, j->t, j->nt, j->idx + j->q0, j->d1 + j->q0, j->d2 + j->q0);
    return NULL;
}

void orbo_knn2_mt(const uint8_t *q, int nq, const uint8_t *t, int nt,
                  int32_t *idx, int32_t *d1, int32_t *d2, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
    knn_job *jobs = malloc(sizeof(knn_job) * nthreads);
    for (int k = 0; k < nthreads; k++) {
        int q0 = (int)((long)nq * k / nthreads), q1 = (int)((long)nq * (k + 1) / nthreads);
        jobs[k] = (knn_job){q, t, q0, q1, nt, idx, d1, d2};
        pthread_create(&th[k], NULL, knn_worker, &jobs[k]);
    }
    for (int k = 0; k < nthreads; k++) pthread_join(th[k], NULL);
    free(th); free(jobs);
}
