// orbx_internal.h -- shared host/device structures of liborbx (not part of the ABI).
#pragma once
#include <stdint.h>
#include <stddef.h>

#define ORBX_MAXL 16
#define ORBX_EDGE 19           // EDGE_THRESHOLD, orbextractor.cpp:135
#define ORBX_MINB 16           // EDGE_THRESHOLD-3, orbextractor.cpp:914
#define ORBX_CELL_W 30.0f      // W, orbextractor.cpp:910
#define ORBX_MAX_STRIPS 8      // nIni supported by the node packing (3 bits)
#define ORBX_MAX_NODES 65534   // per-level node list capacity: 16-bit order arrays in k_octree.  The working limit is the octree's
                               // shared memory: 13 bytes per node + 12 per strip row <= 200 KB

// Per-level geometry.  Everything here is a pure function of (config, image size) and is
// computed once on the host with the reference's exact float32 expressions.
struct OrbxLevel {
    int w, h;            // level size, orbextractor.cpp:659
    int pitch;           // bytes between rows of this level in the pyramid slab (multiple of 128)
    int off;             // byte offset of the level inside one frame's slab
    // gridded FAST, orbextractor.cpp:914-928
    int nCols, nRows, wCell, hCell;
    int cellMagic;       // 65536 / wCell + 1: x / wCell == x * cellMagic >> 16 for x < 1024
    int segBase, nSegs;  // FAST segments (runs of cells in one cell row) of this level inside the segment table
    int winH;            // rows of the FAST window box (hCell + 6): height of this level's TMA box
    // DistributeOctTree, orbextractor.cpp:684-699
    int W, H;            // maxX-minX, maxY-minY
    int nIni, hX, quota;
    int rowBase;         // first entry of this level in the per-frame row-summary arrays
    int slotBase, slotCap; // keypoint slots of this level inside a frame
    // post-processing, orbextractor.cpp:978, :631-637
    float sf, invSf;
    int kpSize;          // 31 * (int)sf
    // resize coefficient tables (entries of OrbxRTab), level l built from level l-1
    int xtabOff, ytabOff;
};

struct OrbxLayout {
    int nlevels;
    int rowsPerFrame;    // sum over levels of nIni*H
    int slotsPerFrame;   // sum over levels of slotCap
    int kpStride;        // records per frame in the output arrays (>= slotsPerFrame)
    int iniTh, minTh;
    int tieRule;
    int totalSegs;
    long long slab;      // pyramid bytes per frame
    OrbxLevel lv[ORBX_MAXL];
};

// One FAST segment: a run of up to ORBX_SEG_W tested columns of horizontally adjacent cells of one cell row
// (orbextractor.cpp:930-947).  The tested pixels of the cells tile the run without gaps: cell j of the run
// covers tested columns [j*wCell, (j+1)*wCell), the last one clipped by the level border.
#define ORBX_SEG_W 112
#define ORBX_FAST_THREADS 128     // threads of a k_fast_segs CTA
#define ORBX_FAST_PITCH 144       // bytes per row of its window and score map: 36 words, so vertical neighbours sit 4 banks apart
struct OrbxSeg {
    uint16_t x0, y0;     // window origin in level coordinates (iniX of the first cell, iniY)
    uint16_t wT;         // tested columns of the run (sum over its cells of window width - 6)
    uint8_t hT;          // tested rows (window height - 6)
    uint8_t level;
    uint16_t ci, cj0;    // cell row / first cell column: (ci*nCols + cj) is a cell's position in the reference's emission order
    uint32_t mQ;         // bits 0-23: 2^20/nQ + 1, nQ = aligned 4-pixel groups covering one tested row of the run;
                         // bits 24-31: rows per band, ceil(hT / (ORBX_FAST_THREADS / nQ))
};

#define ORBX_BLUR_ROWS 36       // output rows per warp band of k_blur; a tile is 4 bands
// one blur tile: 32 words (128 px) x 4 bands of ORBX_BLUR_ROWS rows; x0 in 4-px words, y0 in rows
struct OrbxTile { uint16_t x0, y0; uint8_t level, pad[3]; };

// bilinear coefficient entry, SURVEY A.1.  x axis: {sx0, sx1, c0 | c1 << 16, 0}; y axis: {sy0, sy1, b0, b1}
struct OrbxRTab { int32_t a, b, c, d; };

struct OrbxDbgCand { int32_t xy; int32_t score; };
