// CanvasPartition (wavelets) on the device — shared declarations.
//
// Reference: Src/Canvas/CanvasPartition/{WaveletsRunner,WaveletSegmentation,Segmentation}.cs.
// Pipeline (all on ctx->stream, no host round trip until the results are copied back):
//   1. genome-wide scalars: coverage variability (CV), factor-of-three CMADs, evenness score,
//      per-chromosome median / MAD  -> per-chromosome threshold sigma
//   2. per-chromosome prefix sums of the coverage
//   3. Unbalanced-Haar decomposition: one persistent kernel, dynamic node queue (uh_decompose_kernel)
//   4. per chromosome: hard threshold, reconstruction on the surviving nodes, healing of bad
//      splits, germline refinement (uh_finish_kernel)
#pragma once
#include "common.cuh"
#include "select.cuh"

constexpr int WV_MAX_CHROM = 256;
constexpr int WV_F3_LEVELS = 8;        // FactorOfThreeCoverageVariabilities maxExponent
constexpr int WV_WINDOW_IQR = 10000;   // Segmentation.cs:263,313
constexpr int WV_SCAN_TILE = 2048;

// ---- decomposition tiers -------------------------------------------------------------------
constexpr int UH_SMALL_MAX = 1024;  // nodes up to this size: a whole subtree is done by one warp
constexpr int UH_TINY_MAX = 16;     // nodes up to this size: sequential reference recurrence per thread
constexpr int UH_CHUNK = 8192;      // split positions per chunk ticket of a big node
constexpr int UH_THREADS = 512;
constexpr int UH_QCAP = 1 << 16;    // ticket ring capacity

// Segment kinds of the partition select table (see wavelet.cu)
struct WvSegTable {
    int n_w10, n_w100, n_chrom;
    int base_w10, base_w100, base_chrom, base_f3, base_ev10, base_ev100, base_r10, base_r100;
    int nseg;
};

struct UhNode {  // big node record, stored in the ring slot of its first ticket
    int c, s, e, level;
    int nchunks, done;
    unsigned long long pos;  // ring position of ticket 0
};

struct UhSmallTask {
    int c, s, e, level;
};

struct UhCand {  // node whose coefficient may survive the hard threshold
    unsigned long long key;  // (chrom << 56) | (level << 32) | start  -> sort key
    int s, b, e, level;      // 0-based: left part [s, b], right part [b+1, e]
    int c, pad;
    double coef;
};

struct WvCtl {
    // queues
    unsigned long long q_head, q_tail;
    int bn_count_unused, small_head, small_tail, outstanding;
    int big_done, overflow, cand_count, pad0;
    // scalars
    int cv_has_value, evenness_ok;
    double cv, evenness;
    double f3[WV_F3_LEVELS + 1];
    unsigned ev10_valid, ev100_valid;
    double total_launch_dummy;
    // statistics (bench / DESIGN): bin visits of the decomposition and node counts per tier
    unsigned long long visits_big, visits_small, visits_tiny;
    unsigned long long nodes_big, nodes_small, nodes_tiny;
};
