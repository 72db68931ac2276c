"""BASELINE.json's configurations at their full sizes (3.1 M bins per sample): direct comparison with the oracle where
it finishes in seconds, and size-independent properties — the fused call equals the two separate calls through the
.cleaned round trip, the union of chromosome shards equals the whole, a second run is bit-identical, the pedigree
merge keeps exactly the bins every sample kept."""
import numpy as np
import pytest

from canvas_b200 import synth, textcodec
from oracle import pyoracle as ora

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def wgs():
    return synth.make_sample(config=2, sample=0, scale=1.0)


def _fused(engine, s, **kw):
    return engine.clean_partition_wavelet(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc, **kw)


def test_config2_fused_call_against_oracle_at_full_size(engine, wgs):
    s = wgs
    r = _fused(engine, s, is_germline=True)
    o = ora.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc)
    assert np.array_equal(r["kept_index"], o["kept_index"])
    assert np.array_equal(r["count"].view(np.uint32), o["count"].view(np.uint32))  # normalised counts: bit-exact floats
    assert r["local_sd"] == o["local_sd"]
    off = synth.chrom_offsets(s.chrom[o["kept_index"]], len(s.names))
    p = ora.partition_wavelet(off, ora.f2_roundtrip(o["count"]), is_germline=True, n_threads=16)
    for c, (a, b) in enumerate(zip(r["breakpoints"], p["breakpoints"])):
        assert a.tolist() == b.tolist(), c
    assert r["cv"] == p["cv"] and np.array_equal(r["factor_of_three"], p["factor_of_three"])
    assert abs(r["evenness"] - p["evenness"]) <= 1e-9 * abs(p["evenness"])
    # second run: bit-identical (no run-to-run nondeterminism from atomics or the CUDA graph replay)
    r2 = _fused(engine, s, is_germline=True)
    assert np.array_equal(r2["kept_index"], r["kept_index"]) and np.array_equal(r2["count"].view(np.uint32), r["count"].view(np.uint32))
    assert all(a.tolist() == b.tolist() for a, b in zip(r2["breakpoints"], r["breakpoints"])) and r2["evenness"] == r["evenness"]


def test_config2_fused_equals_two_calls_and_shards(engine, wgs):
    s = wgs
    r = _fused(engine, s, is_germline=True)
    c = engine.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc)
    assert np.array_equal(c["kept_index"], r["kept_index"]) and np.array_equal(c["count"].view(np.uint32), r["count"].view(np.uint32))
    off = synth.chrom_offsets(s.chrom[c["kept_index"]], len(s.names))
    assert np.array_equal(off, r["chrom_off"])
    cov = textcodec.f2_roundtrip(c["count"])
    p = engine.partition_wavelet(off, cov, is_germline=True)
    assert all(a.tolist() == b.tolist() for a, b in zip(p["breakpoints"], r["breakpoints"]))
    assert p["cv"] == r["cv"] and p["evenness"] == r["evenness"]
    # union of three chromosome shards (the multi-GPU decomposition) = the whole
    nc = len(s.names)
    owner = np.arange(nc) % 3
    for k in range(3):
        part = _fused(engine, s, is_germline=True, chrom_selected=(owner == k).astype(np.uint8))
        for ci in range(nc):
            want = r["breakpoints"][ci].tolist() if owner[ci] == k else []
            assert part["breakpoints"][ci].tolist() == want, (k, ci)
        assert part["cv"] == r["cv"]


def test_config3_somatic_pair(engine):
    # tumour / normal: CanvasClean + somatic wavelets (no -g), each sample on its own, half size against the oracle
    for sample, tumour in ((0, True), (1, False)):
        s = synth.make_sample(config=3, sample=sample, scale=0.5, n_events=150, tumour=tumour)
        r = _fused(engine, s, is_germline=False, evenness_window=50000)
        o = ora.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc)
        assert np.array_equal(r["kept_index"], o["kept_index"])
        assert np.array_equal(r["count"].view(np.uint32), o["count"].view(np.uint32))
        off = synth.chrom_offsets(s.chrom[o["kept_index"]], len(s.names))
        p = ora.partition_wavelet(off, ora.f2_roundtrip(o["count"]), is_germline=False, evenness_window=50000, n_threads=16)
        assert all(a.tolist() == b.tolist() for a, b in zip(r["breakpoints"], p["breakpoints"]))
        assert r["cv"] == p["cv"]


def test_config4_trio_clean_merge_hmm(engine):
    # SmallPedigree: three samples cleaned, bins common to all kept (NormalizeCanvasClean), PerSampleHMM per sample
    samples = [synth.make_sample(config=4, sample=k, scale=1.0, n_events=60) for k in range(3)]
    cleaned = []
    # a pedigree shares one bin layout; the synthetic oversized bins push the generator's own coordinates past int32
    # at full scale, so the file coordinates are rebuilt from the bin index (1 kb bins)
    pos = (np.arange(len(samples[0])) - synth.chrom_offsets(samples[0].chrom, len(samples[0].names))[samples[0].chrom]) * 1000
    cleaned_index = []
    for s in samples:
        c = engine.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc)
        k = c["kept_index"]
        cleaned_index.append(k)
        cleaned.append((s.chrom[k], pos[k].astype(np.int32), (pos[k] + 1000).astype(np.int32), c["count"]))
    m = engine.merge_common_bins(cleaned)
    keys = [set(zip(ch.tolist(), st.tolist())) for ch, st, _, _ in cleaned]
    common = keys[0] & keys[1] & keys[2]
    ch0, st0 = cleaned[0][0][m["kept_index"]], cleaned[0][1][m["kept_index"]]
    assert len(m["kept_index"]) == len(common) and set(zip(ch0.tolist(), st0.tolist())) == common
    assert np.all(np.diff(m["kept_index"]) > 0)
    for k in range(3):  # every sample's own count travelled with its bin
        lut = dict(zip(zip(cleaned[k][0].tolist(), cleaned[k][1].tolist()), cleaned[k][3].tolist()))
        pick = np.arange(0, len(ch0), 9973)
        assert [lut[(int(ch0[i]), int(st0[i]))] for i in pick] == m["count"][k][pick].tolist()
    off = synth.chrom_offsets(ch0, len(samples[0].names))
    # the chain bench.py's config 4 times (canvas_b200/pedigree.py: Clean -> merge -> cg_partition_hmm_counts with the
    # float.ToString() round trip on the device) against the oracle fed with the host's emulation of that text
    from canvas_b200 import pedigree
    engine.comm_init(1, 0)
    chain = pedigree.trio_segments(engine, samples)
    assert chain["n_common"] == len(common) and np.array_equal(chain["chrom_off"], off)
    # the index-keyed merge of the chain selects the bins the coordinate-keyed merge selects
    assert np.array_equal(chain["common_index"], cleaned_index[0][m["kept_index"]])
    for k in range(3):
        cov = textcodec.float_default_roundtrip(m["count"][k])  # the merged file prints float.ToString()
        got = engine.partition_hmm(off, cov, per_sample=True)
        want = ora.partition_hmm(off, cov, per_sample=True, n_threads=16)
        assert np.array_equal(got["states"], want["states"])
        assert all(a.tolist() == b.tolist() for a, b in zip(got["breakpoints"], want["breakpoints"]))
        assert all(a.tolist() == b.tolist() for a, b in zip(chain["breakpoints"][k], want["breakpoints"]))
    # the same chain as ONE device-resident call (cg_pedigree_hmm): cleaned lists never leave the GPU between the stages
    s0 = samples[0]
    one = engine.pedigree_hmm(s0.chrom, s0.is_autosome, s0.is_chr_y, s0.start, s0.stop, [s.count for s in samples], s0.gc)
    assert one["n_common"] == len(common) and np.array_equal(one["chrom_off"], off)
    assert np.array_equal(one["common_index"], chain["common_index"])
    assert np.array_equal(one["count"].view(np.uint32), m["count"].view(np.uint32))
    assert one["n_kept"].tolist() == [len(k) for k in cleaned_index]
    for k in range(3):
        assert all(a.tolist() == b.tolist() for a, b in zip(one["breakpoints"][k], chain["breakpoints"][k]))


def test_config2_cbs_whole_genome(engine):
    # CanvasPartition -m CBS on the whole cleaned config-2 sample (3.0 M bins, ~34 000 permutations): identical segments, means
    # and random-stream consumption as the oracle's DNAcopy restatement on every chromosome
    s = synth.make_sample(config=2, sample=0)
    c = engine.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc)
    off = synth.chrom_offsets(s.chrom[c["kept_index"]], len(s.names))
    cov = textcodec.f2_roundtrip(c["count"])
    got = engine.partition_cbs(off, cov)
    want = ora.partition_cbs(off, cov, n_threads=16)
    for ci, (w, g) in enumerate(zip(want["segments"], got["segments"])):
        assert np.array_equal(w["len"], g["len"]) and np.array_equal(w["mean"], g["mean"]), ci
    for k in ("tests", "perms", "perm_steps", "edge_steps"):
        assert want[k] == got[k], k
    assert got["perms"] > 10000


def test_config2_loess_mode_full_size(engine, golden_dir):
    # CanvasClean -m LOESS on the whole config-2 sample.  The oracle needs two minutes for it (its golden-section search
    # evaluates O(n) fits per step), so its full-size result is a committed fixture (tools/make_golden_loess_full.py): the
    # kept bins as a SHA-256, every 997th normalised count, the local-SD metric.  Kept bins exact, values within the
    # north star's 1e-5 (sums of ~3 M terms are grouped differently on the device; the oracle itself is pinned by
    # TestLoessInterpolator)
    import hashlib
    import json
    import os
    g = json.load(open(os.path.join(golden_dir, "loess_config2_full.json")))
    s = synth.make_sample(config=g["config"], sample=g["sample"])
    assert len(s) == g["bins"]
    a = engine.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc, gc_mode=1)
    assert len(a["kept_index"]) == g["kept"]
    assert hashlib.sha256(np.ascontiguousarray(a["kept_index"], np.int32).tobytes()).hexdigest() == g["kept_sha256"]
    assert np.allclose(a["count"][::g["step"]], np.array(g["counts"], np.float32), rtol=1e-5, atol=0)
    assert abs(a["local_sd"] - g["local_sd"]) <= 1e-5 * abs(g["local_sd"])
