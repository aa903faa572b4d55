// orbx_kernels.h -- launchers of the extractor kernels (internal).
#pragma once
#include "orbx_internal.h"
#include <cuda.h>
#include <cuda_runtime.h>

// one 3-D (x, y, frame) tensor map per pyramid level; the array lives in global memory
struct OrbxTensorMaps { CUtensorMap m[ORBX_MAXL]; };

struct orbx_keypoint_pod { float x, y, size, angle, response; int32_t octave, class_id; };

void launch_copy_level0(const uint8_t *src, size_t frameStride, size_t srcPitch, uint8_t *pyr,
                        const OrbxLayout &L, int batch, cudaStream_t st);
void launch_resize(const OrbxTensorMaps &srcMaps, int f0, uint8_t *pyr, const OrbxLayout &L, int level, const int4 *tabs, int batch, int nSM, cudaStream_t st);
void launch_blur(const OrbxTensorMaps &maps, uint8_t *blur, const OrbxLayout &L, const OrbxTile *tiles, int nTiles,
                 const int taps[7], int f0, int batch, cudaStream_t st);
size_t fast_smem_bytes(int winRows, int listCap);
cudaError_t launch_fast(const OrbxTensorMaps &maps, int f0, const OrbxLayout &L, const OrbxSeg *segs, int segBegin, int segCount,
                        uint32_t *cnt, unsigned long long *best, OrbxDbgCand *dbg, int *dbgCount, int dbgCap,
                        int winRows, int listCap, int batch, cudaStream_t st);
size_t octree_smem_bytes(int maxRows, int maxNodes);
cudaError_t launch_octree(const OrbxLayout &L, const uint32_t *cnt, const unsigned long long *best, int2 *slots,
                          int *lvlCount, int maxRows, int maxNodes, int batch, cudaStream_t st);
void launch_describe(const OrbxTensorMaps &mapsA, const OrbxTensorMaps &mapsB, int f0, const OrbxLayout &L, const int2 *slots,
                     const int *lvlCount, const int umax[16], orbx_keypoint_pod *kps, uint8_t *desc, int *counts,
                     int batch, cudaStream_t st);
cudaError_t launch_filter_keypoints(orbx_keypoint_pod *kps, uint8_t *desc, int *counts, int kpStride, int frame0, int nFrames,
                                    const float box[4], cudaStream_t st);
cudaError_t launch_stereo(const OrbxLayout &L, const uint8_t *pyrL, const uint8_t *pyrR, const orbx_keypoint_pod *kl,
                          const uint8_t *dl, const int *nl, const orbx_keypoint_pod *kr, const uint8_t *dr, const int *nr,
                          int nPairs, int frameStep, float mbf, float maxD, float *uRight, float *depth, int *sad, int *nMatches, cudaStream_t st);
