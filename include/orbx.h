/*
 * orbx.h -- C ABI of liborbx.so: the B200 (sm_100a) ORB front end.
 *
 * This is the drop-in boundary for the one hot path this library replaces in
 * chalmers-revere/opendlv-perception-vision-orbslam2 (paths below are relative to
 * the reference tree):
 *
 *   OrbExtractor::OrbExtractor / ExtractFeatures / getters / m_vImagePyramid
 *       include/orbextractor.hpp:90-109, src/orbextractor.cpp:476-642
 *   ORBmatcher::DescriptorDistance and the best / second-best loop around it
 *       include/orbmatcher.hpp:48, src/orbmatcher.cpp:1662-1677, :208-232
 *
 * Plain pointers and sizes only; no C++, OpenCV, torch or CUDA types appear in any
 * signature (a CUDA stream is passed as void*).  All state lives in opaque per-instance
 * handles that own their CUDA stream and device/pinned arenas, so -- like the reference,
 * whose stereo path runs two OrbExtractor instances in two threads
 * (src/orbframe.cpp:73-76) -- different handles may be used concurrently from different
 * host threads; one handle must not be entered by two threads at once.
 *
 * There is no CPU fallback: every entry point that computes runs CUDA kernels and
 * returns ORBX_ERR_CUDA when no sm_100 device is usable.
 *
 * The C++ adapter that keeps the reference's class signatures on top of this ABI is
 * cpp/orbextractor_b200.hpp; the binding notes are in INTEGRATION.md.
 */
#ifndef ORBX_H
#define ORBX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORBX_MAX_LEVELS 16

/* status codes (the reference returns void and prints; see SURVEY.md 8b "Error conventions") */
#define ORBX_OK 0
#define ORBX_ERR_ARG (-1)         /* null pointer, non-positive size, batch > max_batch ...          */
#define ORBX_ERR_SHAPE (-2)       /* image shape the reference itself cannot process (see below)     */
#define ORBX_ERR_CAPACITY (-3)    /* caller's keypoint buffer too small                              */
#define ORBX_ERR_CUDA (-4)        /* CUDA runtime error; text via orbx_last_error                    */
#define ORBX_ERR_NOMEM (-5)

/* 28-byte record, layout-compatible with cv::KeyPoint (pt.x, pt.y, size, angle, response,
 * octave, class_id) as filled by orbextractor.cpp:961-988 and :631-639. */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orbx_keypoint;

/* Replaces the five constructor arguments of OrbExtractor (orbextractor.cpp:476) plus the
 * capacity knobs a device arena needs. */
typedef struct {
    int nfeatures;        /* ORBextractor.nFeatures   (tracking.cpp:104) */
    float scale_factor;   /* ORBextractor.scaleFactor (tracking.cpp:105); supported range (1, 1.35], ORB-SLAM2 settings use 1.2 */
    int nlevels;          /* ORBextractor.nLevels     (tracking.cpp:106), 1..ORBX_MAX_LEVELS */
    int ini_th_fast;      /* ORBextractor.iniThFAST   (tracking.cpp:107) */
    int min_th_fast;      /* ORBextractor.minThFAST   (tracking.cpp:108); 1..254, <= ini_th_fast */
    int max_width;        /* largest image width  this handle will be given */
    int max_height;       /* largest image height this handle will be given */
    int max_batch;        /* frames per orbx_extract_batch call (>=1)       */
    int device;           /* CUDA device ordinal                            */
    int blur_taps[7];     /* 7-tap integer Gaussian (sum ~256); all zero -> {18,34,48,56,48,34,18},
                             the taps of the executable oracle cv2 4.13             */
    int tie_rule;         /* DistributeOctTree equal-size tie (orbextractor.cpp:825 sorts by heap
                             address): 0 = newest node first (canonical), 1 = oldest first          */
} orbx_config;

typedef struct orbx_extractor orbx_extractor;

/* OrbExtractor::OrbExtractor, orbextractor.cpp:476-548 */
int orbx_create(const orbx_config *cfg, orbx_extractor **out);
/* OrbExtractor::~OrbExtractor, orbextractor.cpp:552 */
void orbx_destroy(orbx_extractor *h);
/* text of the last error on this handle (never NULL) */
const char *orbx_last_error(const orbx_extractor *h);

/* OrbExtractor::ExtractFeatures, orbextractor.cpp:582-642: one 8-bit single-channel image
 * (rows `pitch` bytes apart) -> up to kp_cap keypoints and kp_cap x 32 descriptor bytes,
 * *n_out = number written.  kp_cap >= orbx_max_keypoints(h) always suffices. */
int orbx_extract(orbx_extractor *h, const uint8_t *img, int width, int height, size_t pitch,
                 orbx_keypoint *kps, int kp_cap, uint8_t *desc, int *n_out);

/* The same for `batch` independent images of one size (stereo pairs / multi-frame batches;
 * the reference runs these as separate ExtractFeatures calls, orbframe.cpp:73-76).
 * Frame f writes kps[f*kp_cap ...], desc[f*kp_cap*32 ...], n_out[f]. */
int orbx_extract_batch(orbx_extractor *h, const uint8_t *const *imgs, int batch, int width, int height,
                       size_t pitch, orbx_keypoint *kps, int kp_cap, uint8_t *desc, int *n_out);

/* Asynchronous form of orbx_extract_batch: the call only enqueues the work (H2D copies, kernels, D2H copies) and returns a
 * ticket; orbx_wait(ticket) blocks until the results of that call are in kps / desc / n_out, which, like the images, must
 * stay valid and untouched until then.  Up to TWO calls may be in flight on a handle: submit call k+1, then wait for call
 * k -- the H2D copies of call k+1 then run while the last kernels and D2H copies of call k drain, which is where a
 * synchronous call leaves the copy engine idle (the reference consumes the extractors' output the same way: both run in
 * threads, the frame is assembled afterwards, orbframe.cpp:73-78).  A third submit completes the oldest ticket itself.
 * Pinned (page-locked) image and result buffers are DMA'd directly; pageable ones go through the handle's staging.
 * orbx_extract_batch is exactly orbx_extract_batch_async + orbx_wait. */
int orbx_extract_batch_async(orbx_extractor *h, const uint8_t *const *imgs, int batch, int width, int height,
                             size_t pitch, orbx_keypoint *kps, int kp_cap, uint8_t *desc, int *n_out, int *ticket);
/* ORBX_OK, or ORBX_ERR_CAPACITY when a frame produced more than kp_cap keypoints, ORBX_ERR_ARG for an unknown ticket */
int orbx_wait(orbx_extractor *h, int ticket);

/* ---- several GPUs behind one handle: frames are independent, so a batch is cut into contiguous blocks,
 * device slot g taking frames [batch*g/G, batch*(g+1)/G); each slot owns an ordinary orbx_extractor on its GPU and a host
 * thread that submits the slot's block, so the copies and kernels of all GPUs run side by side.  No data-path collective.
 * cfg->max_batch is the largest batch of a call (all slots together), cfg->device is ignored; devices[] lists the CUDA
 * ordinals (an ordinal may appear more than once).  The entry points mirror the single-GPU ones, results land in the
 * caller's arrays exactly where orbx_extract_batch would put them; up to two asynchronous calls may be in flight. */
typedef struct orbx_multi orbx_multi;
int orbx_multi_create(const orbx_config *cfg, const int *devices, int n_devices, orbx_multi **out);
void orbx_multi_destroy(orbx_multi *m);
const char *orbx_multi_last_error(const orbx_multi *m);
int orbx_multi_devices(const orbx_multi *m);
int orbx_multi_max_keypoints(const orbx_multi *m);
int orbx_multi_extract_batch(orbx_multi *m, const uint8_t *const *imgs, int batch, int width, int height, size_t pitch,
                             orbx_keypoint *kps, int kp_cap, uint8_t *desc, int *n_out);
int orbx_multi_extract_batch_async(orbx_multi *m, const uint8_t *const *imgs, int batch, int width, int height, size_t pitch,
                                   orbx_keypoint *kps, int kp_cap, uint8_t *desc, int *n_out, int *ticket);
int orbx_multi_wait(orbx_multi *m, int ticket);
/* the extractor of a slot (device-resident results, stereo matching, pyramid levels of the frames that slot extracted) and
 * the frames of a batch it is handed */
orbx_extractor *orbx_multi_handle(orbx_multi *m, int slot);
int orbx_multi_frame_range(const orbx_multi *m, int batch, int slot, int *first, int *count);

/* Device-resident variant: frames already in HBM (frame f at d_imgs + f*frame_stride), results
 * stay in HBM.  Work is enqueued on `stream` (a cudaStream_t, NULL = the handle's own stream)
 * and NOT synchronised.  Result pointers (valid until the next call on this handle):
 * orbx_device_results.  The handle records an event behind the enqueued work; every later entry point
 * of this handle (the next extraction on any stream, orbx_fetch_results, orbx_filter_keypoints,
 * orbx_stereo_match*, orbx_get_level, the debug taps) orders itself behind that event, so library
 * calls never race with the extraction.  A caller that reads orbx_device_results in kernels of its own
 * must launch them on `stream` or synchronise it. */
int orbx_extract_batch_device(orbx_extractor *h, const uint8_t *d_imgs, size_t frame_stride, size_t pitch,
                              int batch, int width, int height, void *stream);
/* d_kps: batch x kp_stride records; d_desc: batch x kp_stride x 32 bytes; d_counts: batch ints */
/* How many parts orbx_extract_batch_device cuts a batch into (each part runs its stages on a stream pair of its own, so one
 * part's latency-bound stages overlap another's): 0 = default (two for batches of 16 frames and more), 1..4.  A pipe sets 1 on its
 * handles -- there the overlap comes from the other batches in flight, and whole-batch launches are 3 % faster. */
int orbx_set_device_split(orbx_extractor *h, int parts);
int orbx_device_results(orbx_extractor *h, const orbx_keypoint **d_kps, const uint8_t **d_desc,
                        const int **d_counts, int *kp_stride);

/* Several device-resident extractions in flight on one GPU (csrc/orbx_pipe.cu): a pipe owns `depth` (1..8) extractor handles of
 * the same configuration and gives consecutive submissions to them in turn, so that the head of one batch (level-0 copy, the
 * dependent resize launches) runs under the tail of the one before (octree, end of the describe grid).  The device-side twin of
 * orbx_extract_batch_async / orbx_wait; same call shape as the reference's "extract in threads, consume later"
 * (orbframe.cpp:73-78).
 *   orbx_pipe_submit  the frames must be ready at the caller's position in `stream` (NULL = the legacy default stream); the work
 *                     runs on the slot's own stream, nothing is waited for.  *ticket = 1, 2, ...
 *   orbx_pipe_join    makes `stream` wait for that submission and returns its result pointers (as orbx_device_results).  They
 *                     stay valid until the slot is submitted to again, i.e. for depth - 1 further submissions; work the
 *                     caller enqueues on them must be in `stream` before the submission that reuses the slot.
 *   orbx_pipe_handle  the extractor that holds the submission (orbx_stereo_match_batch, orbx_filter_keypoints, orbx_get_level,
 *                     orbx_fetch_results of exactly that batch); NULL once the slot has been reused.
 * Like an extractor handle a pipe is entered by one thread at a time; different pipes and handles are independent. */
typedef struct orbx_pipe orbx_pipe;
int orbx_pipe_create(const orbx_config *cfg, int depth, orbx_pipe **out);
void orbx_pipe_destroy(orbx_pipe *p);
const char *orbx_pipe_last_error(const orbx_pipe *p);
int orbx_pipe_depth(const orbx_pipe *p);
int orbx_pipe_submit(orbx_pipe *p, const uint8_t *d_imgs, size_t frame_stride, size_t pitch, int batch, int width, int height,
                     void *stream, int *ticket);
int orbx_pipe_join(orbx_pipe *p, int ticket, void *stream, const orbx_keypoint **d_kps, const uint8_t **d_desc,
                   const int **d_counts, int *kp_stride);
orbx_extractor *orbx_pipe_handle(orbx_pipe *p, int ticket);

/* Copy the results of the last orbx_extract_batch_device call to host arrays (layout as in
 * orbx_extract_batch).  Waits for `stream` (the stream that call was given; NULL = the handle's). */
int orbx_fetch_results(orbx_extractor *h, void *stream, orbx_keypoint *kps, int kp_cap, uint8_t *desc, int *n_out);

/* OrbFrame::FilterKeyPoints (orbframe.cpp:403-445) on the device-resident results of the last extraction, frames
 * frame0 .. frame0 + n_frames - 1 (left and right images are filtered with the same box, :406-441): key points with
 * box[0] < x < box[1] and box[2] < y < box[3] are removed, the others keep their order; key points, descriptors and
 * counts are compacted in HBM, so orbx_fetch_results / orbx_stereo_match see the filtered frame exactly as
 * OrbFrame::CommonSetup leaves it (:161-168).  As in the reference nothing happens unless box[1] > 2 (:405).
 * Enqueued on the handle's stream (ordered after the extraction and before later calls on this handle). */
int orbx_filter_keypoints(orbx_extractor *h, int frame0, int n_frames, const float box[4]);
/* OrbFrame::ComputeStereoMatches (orbframe.cpp:511-705), the immediate consumer of both extractors:
 * for every left keypoint the sub-pixel u coordinate of its match in the right image and the depth
 * mbf/disparity, -1 where there is none.  Works on the device-resident results (keypoints, descriptors,
 * pyramids) of the last extraction of `left` / frame_left and `right` / frame_right -- two handles as in
 * the reference (orbframe.cpp:73-76) or two frames of one batch -- so no pyramid leaves HBM.
 * mbf = stereo baseline * fx, mb = baseline (orbframe.cpp:545-547 use minZ = mb; pass 0 to reproduce the
 * reference's first-frame behaviour).  n_left = left keypoints, n_matches = matches
 * before the median filter. */
int orbx_stereo_match(orbx_extractor *left, int frame_left, orbx_extractor *right, int frame_right, float mbf, float mb,
                      float *u_right, float *depth, int cap, int *n_left, int *n_matches);
/* The same for n_pairs stereo pairs of ONE extracted batch in one launch: pair p = frames (frame_left0 + p*frame_step,
 * frame_right0 + p*frame_step) of the last extraction of h (L R L R ...: 0, 1, 2).  u_right / depth are [n_pairs][cap]
 * with cap >= orbx_max_keypoints(h); n_left[p] / n_matches[p] as above. */
int orbx_stereo_match_batch(orbx_extractor *h, int n_pairs, int frame_left0, int frame_right0, int frame_step, float mbf, float mb,
                            float *u_right, float *depth, int cap, int *n_left, int *n_matches);

/* upper bound of keypoints per frame for this configuration */
int orbx_max_keypoints(const orbx_extractor *h);

/* Page-locked host memory for callers without the CUDA headers (the header-only adapters in cpp/): images and result
 * arrays allocated here are DMA'd directly -- orbx_extract* writes straight into `kps` / `desc` when they are page-locked
 * and kp_cap == orbx_max_keypoints(h), with no staging copy.  (The reference's cv::Mat / std::vector buffers are
 * pageable, orbextractor.cpp:595-609; the adapter keeps its own page-locked scratch and copies once.) */
int orbx_host_alloc(size_t bytes, void **ptr);
void orbx_host_free(void *ptr);

/* Number of kernel launches the last orbx_extract* call enqueued (all chunks / halves); for benchmark accounting. */
int orbx_last_launches(const orbx_extractor *h);

/* m_vImagePyramid[level] of frame `frame` of the last call (orbextractor.hpp:109; read by
 * OrbFrame::ComputeStereoMatches, orbframe.cpp:518,618-641).  Copies the level lazily to a
 * pinned host buffer owned by the handle; pointer valid until the next extract call. */
int orbx_get_level(orbx_extractor *h, int frame, int level, const uint8_t **host_ptr,
                   int *width, int *height, size_t *pitch);

/* getLevels / getScaleFactors / getInverseScaleFactors / getScaleSigmaSquares /
 * getInverseScaleSigmaSquares (orbextractor.cpp:557-579) and the per-level quotas (:512-523).
 * Any pointer may be NULL; arrays need nlevels entries. Returns nlevels. */
int orbx_scale_tables(const orbx_extractor *h, float *scale, float *inv_scale, float *sigma2,
                      float *inv_sigma2, int *quota);

/* Device time of each stage (ms, averaged over `reps` re-runs on the frames of the last call,
 * serialised with CUDA events on the handle's stream): ms[0] pyramid resize chain, ms[1] gridded
 * FAST, ms[2] DistributeOctTree, ms[3] Gaussian blur, ms[4] orientation + descriptors.  n_ms >= 5. */
int orbx_profile_stages(orbx_extractor *h, int reps, float *ms, int n_ms);

/* ---- stage taps for parity tests (host copies of device intermediates of the last call) ---- */
/* blurred level (the GaussianBlur output of orbextractor.cpp:621-622), tightly packed w*h bytes */
int orbx_debug_blurred(orbx_extractor *h, int frame, int level, uint8_t *dst, size_t dst_bytes, int *width, int *height);
/* enable recording of the gridded-FAST candidates (orbextractor.cpp:930-970) for the next calls */
int orbx_debug_enable_candidates(orbx_extractor *h, int enable);
/* candidates of (frame, level): coords relative to (16,16) like vToDistributeKeys; unordered.
 * Returns count (<= cap written) or <0. */
int orbx_debug_candidates(orbx_extractor *h, int frame, int level, int *xs, int *ys, int *score, int cap);

/* ------------------------------------------------------------------------------------------ */
/* Matcher: ORBmatcher::DescriptorDistance (orbmatcher.cpp:1662-1677) evaluated for every     */
/* query x train pair, with the reference's best / second-best bookkeeping                     */
/* (orbmatcher.cpp:208-232: strict '<' updates, start values 256, index -1).                   */
/* ------------------------------------------------------------------------------------------ */
typedef struct orbm_matcher orbm_matcher;

int orbm_create(int device, int max_queries, int max_train, orbm_matcher **out);
void orbm_destroy(orbm_matcher *m);
const char *orbm_last_error(const orbm_matcher *m);

/* host buffers: q = nq x 32 bytes, t = nt x 32 bytes -> idx[nq] (lowest train index attaining
 * the minimum, -1 if nt == 0), d1[nq] (minimum), d2[nq] (second smallest with multiplicity). */
int orbm_knn2(orbm_matcher *m, const uint8_t *q, int nq, const uint8_t *t, int nt,
              int32_t *idx, int32_t *d1, int32_t *d2);
/* upload / replace the resident train set (map-point descriptors) once, then match many
 * query blocks against it */
int orbm_set_train(orbm_matcher *m, const uint8_t *t, int nt);
int orbm_knn2_resident(orbm_matcher *m, const uint8_t *q, int nq, int32_t *idx, int32_t *d1, int32_t *d2);
/* device-resident variant, enqueued on `stream` (NULL = matcher's stream), not synchronised.
 * d_out = nq x {int32 idx, int32 d1, int32 d2, int32 pad} (16-byte records, ready to be the send
 * buffer of the result gather when queries are sharded over GPUs). */
int orbm_knn2_device(orbm_matcher *m, const uint8_t *d_q, int nq, const uint8_t *d_t, int nt,
                     int32_t *d_out, void *stream);
/* ---- query-sharded kNN-2 over the GPUs of one box, the result gather fused into the kernel ----
 * Every rank (one matcher per GPU; ranks may live in one process or in one process per GPU) holds the whole train set and
 * a RESULT WINDOW of nq_total records.  orbm_knn2_sharded matches the rank's block of queries in ONE kernel launch whose
 * last stage stores the 16-byte {idx, d1, d2, pad} records directly into the window of every rank -- its own and, through
 * NVLink peer mappings, the peers' -- and then waits until the records of all ranks have arrived in its own window; no
 * collective library is involved.  In stream order behind the call, orbm_window_records points at all nq_total records
 * (queries in global order) on this rank's GPU.
 *   orbm_window_create      allocates the window of `rank`; ipc_handle (64 bytes, may be NULL) receives its CUDA IPC handle
 *   orbm_window_attach_ipc  maps the window of a peer rank that lives in ANOTHER process (handle from its orbm_window_create)
 *   orbm_window_attach_peer maps the window of a peer rank of THIS process (enables peer access between the two devices)
 *   orbm_knn2_sharded       d_q = this rank's nq_local queries (global indices q_offset ..), d_t = nt train rows, both on the
 *                           matcher's device, 16-byte aligned; enqueued on `stream` (NULL = the matcher's), not synchronised.
 *                           All ranks must make the call (a rank whose peer never arrives gives up after 5 s).
 *   orbm_window_status      synchronises `stream` and reports a peer that did not arrive (ORBX_ERR_CUDA)
 *   orbm_window_fetch       the same, then copies the first nq records of the last call to the host (nq x 4 int32)
 * The window double-buffers by call parity, so a rank may already run call e+1 while a peer still reads the result of call e
 * in kernels enqueued before its own call e+1. */
int orbm_window_create(orbm_matcher *m, int nq_total, int n_ranks, int rank, void *ipc_handle);
int orbm_window_attach_ipc(orbm_matcher *m, int peer_rank, const void *ipc_handle);
int orbm_window_attach_peer(orbm_matcher *m, int peer_rank, orbm_matcher *peer);
int orbm_knn2_sharded(orbm_matcher *m, const uint8_t *d_q, int nq_local, int q_offset, const uint8_t *d_t, int nt, void *stream);
int orbm_window_status(orbm_matcher *m, void *stream);
int orbm_window_fetch(orbm_matcher *m, void *stream, int32_t *records, int nq);
int orbm_window_records(orbm_matcher *m, const int32_t **d_records);
/* The same inside ONE process: a matcher per listed device, windows attached to one another through peer access.
 * orbm_multi_knn2: queries (host) are cut into blocks, every GPU matches its block against its copy of the train set
 * (orbm_multi_set_train) and the kernels exchange the records among themselves; the host reads the complete result from
 * the first GPU's window. */
typedef struct orbm_multi orbm_multi;
int orbm_multi_create(const int *devices, int n_devices, int max_queries, int max_train, orbm_multi **out);
void orbm_multi_destroy(orbm_multi *mm);
const char *orbm_multi_last_error(const orbm_multi *mm);
int orbm_multi_devices(const orbm_multi *mm);
int orbm_multi_set_train(orbm_multi *mm, const uint8_t *t, int nt);
int orbm_multi_knn2(orbm_multi *mm, const uint8_t *q, int nq, int32_t *idx, int32_t *d1, int32_t *d2);
orbm_matcher *orbm_multi_matcher(orbm_multi *mm, int slot);
/* Candidate-list matching: the inner loop of SearchByProjection / SearchByBoW / SearchForTriangulation
 * (orbmatcher.cpp:76-114, :208-232, :1337-1483).  Query i is compared with the train rows
 * indices[offsets[i] .. offsets[i+1]) in that order (CSR); the sequential exclusion rules of the callers
 * stay on the host, which builds the lists.  Per query: idx1/d1 = best, idx2/d2 = second best under the
 * reference's strict '<' updates (earlier list position wins ties); -1 / 256 where absent.  idx2 lets the
 * caller look up the octave the reference tracks as bestLevel2 (orbmatcher.cpp:105-113). */
int orbm_knn2_csr(orbm_matcher *m, const uint8_t *q, int nq, const uint8_t *t, int nt, const int32_t *offsets,
                  const int32_t *indices, int32_t *idx1, int32_t *d1, int32_t *idx2, int32_t *d2);
/* device-resident variant; d_out = nq x {idx1, d1, d2, idx2} int32 records; enqueued on `stream`, not synchronised */
int orbm_knn2_csr_device(orbm_matcher *m, const uint8_t *d_q, int nq, const uint8_t *d_t, const int32_t *d_offsets,
                         const int32_t *d_indices, int32_t *d_out, void *stream);
/* DescriptorDistance of every (query, candidate) entry of the CSR lists, dist[k] for k in [0, offsets[nq]) in list order.
 * For the drivers whose exclusion rules depend on earlier matches of the same call (SearchByProjection: a frame keypoint
 * that already carries a map point is skipped, orbmatcher.cpp:87-89; Fuse; SearchBySim3): the device evaluates all
 * distances at once, the host replays the reference's loop (:76-124) over the precomputed distances in its own order. */
int orbm_distance_csr(orbm_matcher *m, const uint8_t *q, int nq, const uint8_t *t, int nt, const int32_t *offsets,
                      const int32_t *indices, int32_t *dist);
/* One frame as ORBmatcher::SearchByProjection reads it (include/orbframe.hpp:141-175): n key points. */
typedef struct {
    const orbx_keypoint *keys;      /* m_undistortedKeys.data() (cv::KeyPoint layout) */
    const float *u_right;           /* mvuRight */
    const uint8_t *occupied;        /* != 0: m_mapPoints[i] is set and has observations (orbmatcher.cpp:87-89); may be NULL */
    const uint8_t *desc;            /* m_descriptors, n x 32 */
    int32_t n;
    float min_x, min_y, max_x, max_y;   /* OrbFrame::m_minX .. m_maxY (orbframe.cpp:482-503) */
} orbm_frame_view;
/* ORBmatcher::SearchByProjection(frame, map points, th) (orbmatcher.cpp:42-124) for all map points in one pass, the
 * frame side included: OrbFrame::AssignFeaturesToGrid / PosInGrid (orbframe.cpp:192-211, :381-393; 64 x 48 cells, cell
 * lists in key-point order) and OrbFrame::GetFeaturesInArea (:308-380; cells column by column, level window
 * [level-1, level], |dx| < r and |dy| < r) are evaluated on the device with the reference's float arithmetic.
 * Map point i: descriptor mp_desc[i], projection (mp_x, mp_y)[i] (mTrackProjX/Y), mp_level[i] (mnTrackScaleLevel) and
 * mp_radius[i] = r * m_scaleFactors[level] exactly as the caller computes it at :54-67; only map points that pass the
 * caller's GetTrackInView / IsCorrupt tests (:51-55) are handed in.  Candidates are skipped as at :87-96, best / second
 * best follow the strict '<' updates of :102-114, acceptance is :116-123 with TH_HIGH = th_high and mfNNratio = nnratio.
 * The reference's loop is sequential: an accepted map point is stored in m_mapPoints at once (:121) and, if it has
 * observations, :87-89 hides its key point from every later map point of the same call.  mp_observed[i] != 0 says map point
 * i has GetObservingKeyFrameCount() > 0 (NULL: none has -- in Tracking::SearchLocalPoints all have); the device evaluates
 * all map points at once and repeats the pass until no choice changes, which reproduces the sequential result exactly
 * (map point i is final after pass i + 1 at the latest; two or three passes in practice).
 * mp_match[i] = key point accepted for map point i or -1; assigned[k] = the map point the reference's loop leaves in
 * m_mapPoints[k] among those handed in (the last accepted one) or -1; *nmatches = the function's return value. */
int orbm_search_by_projection(orbm_matcher *m, const orbm_frame_view *frame, const uint8_t *mp_desc, const float *mp_x,
                              const float *mp_y, const int32_t *mp_level, const float *mp_radius, const uint8_t *mp_observed,
                              int n_mp, float nnratio, int th_high, int32_t *mp_match, int32_t *assigned, int32_t *nmatches);
/* OrbFrame::AssignFeaturesToGrid (orbframe.cpp:192-211) with PosInGrid (:381-393): the 64 x 48 feature grid of one frame as
 * CSR.  Cell (ix, iy) = m_grid[ix][iy] is cell_items[cell_start[ix * 48 + iy] .. cell_start[ix * 48 + iy + 1]), key-point
 * indices in increasing order exactly as the reference pushes them; key points PosInGrid rejects are in no cell.
 * cell_start has 64 * 48 + 1 entries, cell_items n. */
int orbm_assign_grid(orbm_matcher *m, const orbx_keypoint *keys, int n, float min_x, float min_y, float max_x, float max_y,
                     int32_t *cell_start, int32_t *cell_items);
/* OrbFrame::GetFeaturesInArea (orbframe.cpp:308-380) for nq windows of one frame at once, plus DescriptorDistance of every
 * feature found -- the building block of the window-based drivers whose acceptance is sequential (SearchForInitialization,
 * orbmatcher.cpp:411-528: vMatchedDistance; SearchByProjection(CurrentFrame, LastFrame), :1337-1483; Fuse): the host keeps
 * the projection arithmetic and replays its loop over the lists.  Window i: centre (q_x, q_y)[i], half size q_r[i], levels
 * [q_min_level[i], q_max_level[i]] with the reference's meaning of -1 (:337, :352-364).  offsets[nq + 1] / indices: the
 * features of every window in the order the reference returns them; dist[k] = DescriptorDistance(q_desc[i], frame
 * descriptor indices[k]) when q_desc != NULL.  cap = entries the arrays hold; *n_entries = entries found (ORBX_ERR_CAPACITY
 * when that is more than cap: call again with larger arrays). */
int orbm_area_distances(orbm_matcher *m, const orbm_frame_view *frame, const uint8_t *q_desc, const float *q_x, const float *q_y,
                        const float *q_r, const int32_t *q_min_level, const int32_t *q_max_level, int nq, int32_t *offsets,
                        int32_t *indices, int32_t *dist, int cap, int32_t *n_entries);
/* OrbMapPoint::ComputeDistinctiveDescriptors (orbmappoint.cpp:314-383), batched over map points.
 * Point p observes the rows indices[offsets[p] .. offsets[p+1]) of the descriptor pool desc[n_desc][32] (what the
 * reference gathers from its key frames, :328-337).  Per point: all-pairs DescriptorDistance (:350-358), per row the
 * median = element (int)(0.5*(N-1)) of the sorted row (:367), and best[p] = position inside the point's list of the
 * first row with the least median (:369-373) -- the descriptor the reference clones into m_descriptor (:382);
 * -1 for an empty list (the reference returns early, :322-325, :339-342).  median[p] (optional) = that median.
 * At most 768 observations per point. */
int orbm_distinctive(orbm_matcher *m, const uint8_t *desc, int n_desc, const int32_t *offsets, const int32_t *indices,
                     int n_points, int32_t *best, int32_t *median);
/* DescriptorDistance for n independent pairs a[i], b[i] (32 bytes each) -> out[i] */
int orbm_distance_pairs(orbm_matcher *m, const uint8_t *a, const uint8_t *b, int n, int32_t *out);

/* Measured POPC issue rate of this GPU in lane-operations per clock per SM (denominator of the matcher's
 * INT-pipe roofline: pairs/s = SMs x clock x rate / 8). */
int orbm_measure_popc(orbm_matcher *m, double *popc_per_clk_per_sm);

/* library build info: "orbx <version> sm_100a" */
const char *orbx_version(void);

/* ---- bag-of-words descent ----------------------------------------------------------------
 * OrbVocabulary::transform5 (orbvocabulary.cpp:203-242), the per-feature part of transform4 (:168-201), which
 * OrbFrame::ComputeBoW / OrbKeyFrame::ComputeBoW call once per frame (orbframe.cpp:395-402, orbkeyframe.cpp:61-70).
 * The tree is passed as arrays (what OrbVocabulary's text loader, :39-118, builds in m_nodes): node 0 is the root;
 * the children of node v are child_ids[child_off[v] .. child_off[v+1]) in the order of m_nodes[v].children; a node
 * without children is a leaf and carries word_id[v] (>= 0) and weight[v]; node_desc[v] is its 32-byte descriptor;
 * L = m_L.  The std::map bookkeeping of transform4 (addWeight / addFeature / normalize) stays on the host. */
typedef struct orbv_vocab orbv_vocab;
int orbv_create(int device, int n_nodes, const int32_t *child_off, const int32_t *child_ids, const uint8_t *node_desc,
                const int32_t *word_id, const double *weight, int L, orbv_vocab **out);
void orbv_destroy(orbv_vocab *v);
const char *orbv_last_error(const orbv_vocab *v);
/* n features (rows of 32 bytes) -> word_id[i], weight[i] (optional), node_id[i] (optional) = node passed at level
 * L - levels_up, 0 = root when that level is <= 0 (:211-212).  Strict '<' on the distances: the first child with the
 * least distance is followed (:224-232). */
int orbv_transform(orbv_vocab *v, const uint8_t *desc, int n, int levels_up, int32_t *word_id, double *weight, int32_t *node_id);
/* device-resident variant: descriptors in HBM with `desc_stride` bytes between rows (e.g. the extractor's d_desc),
 * d_word_node = n x {int32 word, int32 node}; enqueued on `stream` (NULL = the handle's), not synchronised */
int orbv_transform_device(orbv_vocab *v, const uint8_t *d_desc, size_t desc_stride, int n, int levels_up,
                          int32_t *d_word_node, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ORBX_H */
