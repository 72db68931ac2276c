"""CanvasBin counting: the oracle and the host pairing logic against the reference's own test
(CanvasTest/TestCanvasBin.cs:14-78) and hand-checked cases.  CPU only."""
import numpy as np
import pytest

from canvas_b200 import binning
from oracle import pyoracle

SAM_PAIRED_PROPER = 0x1 | 0x2


def _pair(pos1, pos2, mq1, mq2):
    flags = binning.flags_from_sam([SAM_PAIRED_PROPER, SAM_PAIRED_PROPER])
    return dict(flags=flags, pos=[pos1, pos2], mate_pos=[pos2, pos1], ref_id=[0, 0], mate_ref_id=[0, 0],
                frag_len=[100, -100], mapq=[mq1, mq2])


@pytest.mark.parametrize("pos1,pos2", [(100, 120), (100, 100)])
@pytest.mark.parametrize("mq1,mq2,expect", [(10, 10, 1), (10, 2, 0), (2, 10, 0), (2, 2, 0)])
def test_bin_one_alignment_reference_cases(pos1, pos2, mq1, mq2, expect):
    # TestCanvasBin.TestBinOneAlignment: one bin chr1:100-200, quality threshold 3
    a = _pair(pos1, pos2, mq1, mq2)
    r = pyoracle.bin_alignments(a["flags"], a["pos"], a["mate_pos"], a["ref_id"], a["mate_ref_id"], a["frag_len"], a["mapq"],
                                [7, 7], 3, [100], [200])
    assert r["count"].tolist() == [expect]
    fs, fe, undo = binning.pair_fragments(a["flags"], a["pos"], a["mate_pos"], a["ref_id"], a["mate_ref_id"], a["frag_len"],
                                          a["mapq"], ["ReadName", "ReadName"], 3)
    # the geometric half, done here in plain Python (the kernel does it on the GPU tier)
    best = [0 if min(200, e) - max(100, s) > 0 else -1 for s, e in zip(fs.tolist(), fe.tolist())]
    count = sum(b >= 0 for b in best) - sum(best[u] >= 0 for u in undo.tolist())
    assert count == expect


def test_find_best_bin_ties_and_gaps():
    # FindBestBin: largest overlap, first bin on ties, stop at the first non-overlapping bin
    bs, be = [0, 100, 200, 400], [100, 200, 300, 500]
    flags = binning.flags_from_sam([SAM_PAIRED_PROPER] * 4)
    pos = [50, 150, 290, 600]
    fl = [100, 100, 105, 50]  # 50/50 tie -> bin 0; 50/50 tie -> bin 1; 10 in bin 2, ends before bin 3; right of all
    r = pyoracle.bin_alignments(flags, pos, [p + 1000 for p in pos], [0] * 4, [0] * 4, fl, [60] * 4, [1, 2, 3, 4], 3, bs, be)
    assert r["count"].tolist() == [1, 1, 1, 0] and r["usable"] == 3


def test_bin_hits_hand_case():
    # 'nn' prefix skipped, bins close at every 2nd possible position, hits capped at 10, trailing partial bin dropped
    bases = b"nnACGTGGCCAT"
    possible = [0, 0, 1, 0, 1, 1, 0, 1, 1, 0, 1, 0]
    hits = [9, 9, 3, 5, 12, 1, 7, 0, 255, 2, 4, 6]
    r = pyoracle.bin_hits(hits, possible, bases, 2)
    assert r["start"].tolist() == [2, 5, 8]
    assert r["stop"].tolist() == [5, 8, 11]
    assert r["count"].tolist() == [3 + 10, 1 + 0, 10 + 4]
    assert r["gc"].tolist() == [int(100 * 2 / 3), int(100 * 2 / 3), int(100 * 2 / 3)]


def test_bin_hits_weighted_rounding():
    bases = b"ACGT" * 4
    possible = [1] * 16
    hits = [1] * 16
    ratio = np.ones(101, np.float32)
    ratio[40] = 2.0
    r = pyoracle.bin_hits(hits, possible, bases, 5, mode=1, read_gc=[40] * 16, obs_vs_exp=ratio)
    assert r["count"].tolist() == [2, 2, 2]  # 2.5 rounds half to even


def test_hit_array_saturates():
    h = binning.hit_array(4, [1] * 300 + [3])
    assert h.tolist() == [0, 255, 0, 1]


def test_bin_screen_by_hand():
    hits = np.array([1, 0, 3, 2, 0, 5, 1, 1], np.uint8)
    possible = np.array([1, 1, 0, 1, 1, 1, 0, 1], bool)
    r = pyoracle.bin_screen(hits, possible, [4], [6])           # the filter interval [4, 6) removes two possible positions
    assert r["possible"].tolist() == [True, True, False, True, False, False, False, True]
    assert r["hits"].tolist() == [1, 0, 0, 2, 0, 0, 0, 1]
    assert (r["observed"], r["n_possible"]) == (3, 4)
    # GetBinSize (CanvasBin.cs:79-83): (int)(countsPerBin / median rate), mean of the middles for an even count
    assert binning.bin_size_from_rates(100, [0.5, 0.25, 0.1]) == 400
    assert binning.bin_size_from_rates(100, [0.5, 0.25]) == 266
    assert binning.bin_size_from_rates(100, [0.0]) == -2 ** 31


def test_read_gc_by_hand():
    # mean fragment 2, cutoff 3: positions 0 .. len - 7 - 1 get a value
    bases = b"GGCCAATTGCATATATAT"
    frag = np.zeros(len(bases), np.int16)
    frag[1] = 4      # own length
    frag[2] = 100    # capped at 3 * mean = 6
    frag[3] = -5     # negative: no base is counted
    r = pyoracle.bin_read_gc(bases, frag, 2, np.ones(len(bases), np.uint8))
    n_set = len(bases) - 2 * 3 - 1
    assert r["read_gc"][:4].tolist() == [100, 75, 33, 0]          # GG | GCCA | CCAATT -> 2/6 | nothing
    assert r["read_gc"][n_set:].tolist() == [0] * (len(bases) - n_set)
    assert r["expected"].sum() == len(bases) and r["observed"].sum() == len(bases)
    assert binning.mean_fragment_size([(1000, 4), (0, 0), (900, 3)]) == 275
    t = binning.observed_vs_expected_gc(r["expected"], 2 * r["expected"])
    assert t.dtype == np.float32 and np.allclose(t[r["expected"] > 0], 1.0) and np.allclose(t[r["expected"] == 0], 0.5)
