// CanvasPartition (wavelets) on the device — shared declarations.
//
// Reference: Src/Canvas/CanvasPartition/{WaveletsRunner,WaveletSegmentation,Segmentation}.cs.
// One call, no host round trip between its stages (the fused Clean + Partition call waits once, for the per-chromosome
// survivor counts of Clean):
//   1. per-chromosome prefix sums of the coverage (main stream; in the fused call before the host has planned)
//   2. genome-wide scalars: per-chromosome / per-window medians and MADs on integer hundredths -> coverage variability (CV),
//      per-chromosome threshold sigma (main stream); factor-of-three CMADs and evenness score (side stream)
//   3. Unbalanced-Haar decomposition, one pipeline per chromosome on its own stream: chains of big nodes (thread-block
//      clusters) -> mid subtrees (one CTA each) -> small subtrees (one warp each) -> tiny subtrees (one thread each)
//   4. per chromosome, as soon as ITS tree is done: hard threshold, reconstruction on the surviving nodes, healing of bad
//      splits, germline refinement (uh_finish_kernel)
// The stream / event / graph structure is drawn in DESIGN.md §3.
#pragma once
#include "common.cuh"
#include "select.cuh"

constexpr int WV_MAX_CHROM = 256;
constexpr int WV_F3_LEVELS = 8;        // FactorOfThreeCoverageVariabilities maxExponent
constexpr int WV_WINDOW_IQR = 10000;   // Segmentation.cs:263,313
constexpr int WV_SCAN_TILE = 2048;

// ---- decomposition tiers -------------------------------------------------------------------
constexpr int UH_MID_MAX = 16384;   // nodes above this size: stage A (thread-block clusters)
constexpr int UH_SMALL_MAX = 1024;  // nodes up to this size: a whole subtree is done by one warp (stage S)
constexpr int UH_TINY_MAX = 16;     // nodes up to this size: sequential reference recurrence per thread (stage T)
constexpr int UH_MID_THREADS = 0;    // mid stage: 0 = every node straight from L2, 256 threads (default); CANVAS_MID_THREADS=256 / 512 / 1024 = subtree staged in shared memory (A/B: profiles/rd2o_*, no gain)
constexpr int UH_SMALL_THREADS = 256;
constexpr int UH_THREADS = 512;
constexpr int UH_CLUSTER = 8;        // CTAs (SMs) cooperating on one chain of big nodes
constexpr int UH_QCAP = 1 << 16;    // total capacity of the ticket rings (every chromosome owns a power-of-two slice)

// Segment kinds of the partition select table (see wavelet.cu)
struct WvSegTable {
    int n_w10, n_w100, n_chrom;
    int base_w10, base_w100, base_chrom, base_f3, base_ev10, base_ev100, base_r10, base_r100;
    int nseg;
};

struct UhBigTask {  // a big node waiting for a chain worker
    int c, s, e, level;   // c < 0: slot empty
    double base, endv;    // prefix sums just before s and at e
};

struct UhTask {  // a node waiting for stage M or S
    int c, s, e, level;
    double base, endv;  // prefix sums just before s and at e
};

struct UhTinyTask {  // a node waiting for stage T (it reads the coverage itself)
    int c, s, e, level;
};

struct UhCand {  // node whose coefficient may survive the hard threshold
    unsigned long long key;  // (chrom << 56) | (level << 32) | start  -> sort key
    int s, b, e, level;      // 0-based: left part [s, b], right part [b+1, e]
    int c, pad;
    double coef;
};

// Every hot word of the queues sits in its own 128-byte line: thousands of idle warps poll
// `big_done` while the big workers run atomics on the ring counters.
struct alignas(128) WvPadU64 { unsigned long long v; unsigned long long pad[15]; };
struct alignas(128) WvPadI32 { int v; int pad[31]; };

// Every chromosome runs its own pipeline (chain -> mid -> small -> tiny -> finish on its own stream), so that its later
// stages and its finish start as soon as ITS tree is ready while other chromosomes are still decomposing.  Queue state and
// list slices are therefore per chromosome.
struct UhChromCtl {
    WvPadU64 q_head_, q_tail_;
    WvPadI32 outstanding_, big_done_;
    WvPadI32 mid_head_, mid_tail_, small_head_, small_tail_, tiny_tail_, cand_count_;
};

struct UhChromPlan {      // slices of the shared arrays, filled by the host from the chromosome lengths
    int ring_base, ring_cap;   // ring_cap is a power of two
    int mid_base, mid_cap;
    int small_base, small_cap;
    int tiny_base, tiny_cap;
    int cand_base, cand_cap;
    int pad[2];
};

struct WvCtl {
    WvPadI32 overflow_;
    // scalars
    int cv_has_value, evenness_ok;
    double cv, evenness;
    double f3[WV_F3_LEVELS + 1];
    unsigned ev10_valid, ev100_valid;
    double total_launch_dummy;
    // statistics (bench / DESIGN): bin visits of the decomposition and node counts per tier
    unsigned long long visits_big, visits_small, visits_tiny;
    unsigned long long nodes_big, nodes_small, nodes_tiny;
    unsigned long long cand_total;
    // timeline of the decomposition kernels (%globaltimer, ns)
    unsigned long long t_first, t_big_done, t_last;
    unsigned long long multi_chunk_nodes, queue_hops;
};
