// ORACLE — test infrastructure only (never linked into or called by the product path).
// CPU restatement of CanvasBin's two counting loops:
//   ora_bin_screen      CanvasBin.cs:668-692 (ExcludeTagsOverlappingFilterFile), :699-716 (ScreenObservedTags), :56-58 with
//                       HitArray.cs:24-32 and CanvasBin.cs:147-157 (the two counts behind a chromosome's rate)
//   ora_bin_read_gc     CanvasBin.cs:450-497 (GC content of the read at every position, counted base by base as the reference does)
//                       and :341-358 (expected / observed read counts per GC bin)
//   ora_bin_hits        CanvasBin.cs:568-661  (BinCountsForChromosome, no predefined bins)
//   ora_bin_alignments  FragmentBinner.cs:256-371 (BinOneAlignment + FindBestBin, sequential, with the
//                       read-name dictionary exactly as the reference keeps it)
// Pinned by CanvasTest/TestCanvasBin.cs:17-78 (tests/test_bin_oracle.py).
#include <cmath>
#include <cstdint>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "oracle.h"

extern "C" int64_t ora_bin_hits(int64_t len, const uint8_t* hits, const uint8_t* possible, const char* bases, int bin_size, int mode,
                                const uint8_t* read_gc, const float* obs_vs_exp, int64_t max_bins, int32_t* start, int32_t* stop,
                                int32_t* count, uint8_t* gc) {
    int64_t pos = 0;
    while (pos < len && bases[pos] == 'n') pos++;  // :582-583 (the reference indexes past the end when all are 'n')
    int64_t nb = 0;
    int bin_start = (int)pos, possible_count = 0, gc_count = 0, nt_count = 0;
    for (; pos < len; pos++) {
        const char b = bases[pos];
        if (b == 'C' || b == 'c' || b == 'G' || b == 'g') gc_count++;  // :587-592
        nt_count++;                                                     // :593 (the upper-case switch never excludes anything)
        if (possible[pos]) possible_count++;                            // :596
        if (possible_count == bin_size) {                               // :611-612
            float weighted = 0.0f;
            int observed = 0;
            for (int64_t p = bin_start; p <= pos; p++) {                // :614-637
                if (!possible[p]) continue;
                if (mode == 0) {
                    observed += hits[p] < 10 ? hits[p] : 10;            // :621-623, Math.Min(10, hits)
                } else {
                    const float w = (float)hits[p] / obs_vs_exp[read_gc[p]];  // :628-633, float arithmetic
                    weighted += w < 10.0f ? w : 10.0f;
                }
            }
            if (mode == 1) observed = (int)std::nearbyint((double)weighted);  // Math.Round: half to even
            const int gcp = (int)(100.0f * (float)gc_count / (float)nt_count);  // :638
            if (nb < max_bins) {
                start[nb] = bin_start;
                stop[nb] = (int)pos + 1;  // :652
                count[nb] = observed;
                gc[nb] = (uint8_t)gcp;
            }
            nb++;
            bin_start = (int)pos + 1;  // :654-659
            possible_count = gc_count = nt_count = 0;
        }
    }
    return nb;
}

// flags: bit0 mapped, bit1 mate mapped, bit2 primary, bit3 paired, bit4 proper pair, bit5 duplicate, bit6 failed QC
extern "C" int64_t ora_bin_alignments(int64_t n, const uint8_t* flags, const int32_t* pos, const int32_t* mate_pos, const int32_t* ref_id,
                                      const int32_t* mate_ref_id, const int32_t* frag_len, const uint32_t* mapq, const int64_t* name_id,
                                      uint32_t quality_threshold, int64_t n_bins, const int32_t* bin_start, const int32_t* bin_stop,
                                      int32_t* count) {
    std::unordered_map<int64_t, int> name_to_bin;
    std::unordered_set<int64_t> same_pos;
    int64_t usable = 0;
    int64_t idx_start = 0;
    for (int64_t b = 0; b < n_bins; b++) count[b] = 0;
    for (int64_t i = 0; i < n; i++) {
        const uint8_t f = flags[i];
        if (!(f & 1) || !(f & 2) || !(f & 4) || !((f & 8) && (f & 16))) continue;  // :259-262
        const bool bad = (f & 32) || (f & 64) || mapq[i] == 255u || mapq[i] < quality_threshold;  // :322-333
        auto it = name_to_bin.find(name_id[i]);
        if (it != name_to_bin.end()) {  // :267-277
            if (bad) { usable--; count[it->second]--; }
            name_to_bin.erase(it);
            continue;
        }
        if (bad) continue;
        if (ref_id[i] != mate_ref_id[i]) continue;
        if (pos[i] > mate_pos[i]) continue;  // IsRightMostInPair
        if (pos[i] == mate_pos[i]) {         // :284-292
            auto sp = same_pos.find(name_id[i]);
            if (sp != same_pos.end()) { same_pos.erase(sp); continue; }
            same_pos.insert(name_id[i]);
        }
        if (frag_len[i] == 0) continue;
        const int fs = pos[i], fe = pos[i] + frag_len[i];
        while (idx_start < n_bins && bin_stop[idx_start] <= fs) idx_start++;
        if (idx_start >= n_bins) continue;
        int best = -1, best_ov = 0;  // FindBestBin :353-371
        for (int64_t b = idx_start; b < n_bins; b++) {
            const int os = bin_start[b] > fs ? bin_start[b] : fs;
            const int oe = bin_stop[b] < fe ? bin_stop[b] : fe;
            const int ov = oe - os;
            if (ov <= 0) break;
            if (ov > best_ov) { best_ov = ov; best = (int)b; }
        }
        if (best >= 0) { usable++; count[best]++; name_to_bin[name_id[i]] = best; }
    }
    return usable;
}

extern "C" void ora_bin_screen(int64_t len, uint8_t* hits, uint8_t* possible, int64_t n_filter, const int32_t* filter_start,
                               const int32_t* filter_stop, int64_t* n_observed, int64_t* n_possible) {
    for (int64_t k = 0; k < n_filter; k++)
        for (int64_t i = filter_start[k]; i < filter_stop[k]; i++) possible[i] = 0;
    for (int64_t i = 0; i < len; i++)
        if (!possible[i]) hits[i] = 0;
    int64_t obs = 0, pos = 0;
    for (int64_t i = 0; i < len; i++) {
        if (hits[i] > 0) obs++;
        if (possible[i]) pos++;
    }
    *n_observed = obs;
    *n_possible = pos;
}

extern "C" void ora_bin_read_gc(int64_t len, const char* bases, const int16_t* frag_len, int mean_frag, const uint8_t* hits, uint8_t* read_gc,
                                int64_t* expected, int64_t* observed) {
    const int cutoff = 3;
    for (int64_t i = 0; i < len; i++) read_gc[i] = 0;
    uint32_t gc_counter = 0;
    for (int64_t pos = 0; pos < len - (int64_t)mean_frag * cutoff - 1; pos++) {
        int16_t current = 0;
        if (frag_len[pos] == 0) current = (int16_t)mean_frag;
        else current = (int16_t)std::min<int>(frag_len[pos], mean_frag * cutoff);
        for (int64_t i = pos; i < pos + current; i++) {
            switch (bases[i]) {
                case 'C': case 'c': case 'G': case 'g': gc_counter++; break;
                default: break;
            }
        }
        const int64_t v = (int64_t)100 * (int64_t)gc_counter / (int64_t)current;
        read_gc[pos] = (uint8_t)std::min<int64_t>(v, 101);
        gc_counter = 0;
    }
    for (int64_t i = 0; i < len; i++) {
        expected[read_gc[i]]++;
        observed[read_gc[i]] += hits[i];
    }
}
