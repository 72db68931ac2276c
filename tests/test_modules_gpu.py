"""The module mains end to end on files (GPU): same command lines and file formats as the reference's
CanvasClean / CanvasPartition, checked against the oracle driven through the same host code."""
import gzip
import os

import numpy as np
import pytest

from canvas_b200 import fileio, modules, synth, textcodec
from oracle import pyoracle as ora

pytestmark = pytest.mark.gpu


def test_clean_then_partition_on_files(tmp_path, engine):
    modules._engine = engine
    s = synth.make_sample(config=2, sample=11, scale=0.07, n_events=120)
    binned = str(tmp_path / "s.binned")
    cleaned = str(tmp_path / "s.cleaned")
    part = str(tmp_path / "s.partitioned")
    lsd = str(tmp_path / "LocalSdMetric.txt")
    evn = str(tmp_path / "EvennessMetric.txt")
    vaf = str(tmp_path / "s.SNV.txt.gz")
    gzip.open(vaf, "wt").write("")
    cfg = str(tmp_path / "CanvasPartitionParameters.json")
    open(cfg, "w").write('{"EvennessScoreWindow": 6000}')
    fileio.write_binned(binned, s.names, s.chrom, s.start, s.stop, s.count, s.gc)
    assert modules.main(["CanvasClean", "-i", binned, "-o", cleaned, "-g", "-s", "-r",
                         f"--local-sd-metric-file={lsd}"]) == 0
    # expected .cleaned: oracle on what the reader parses back from the .binned text
    sb = fileio.read_binned(binned)
    o = ora.clean(sb.chrom, sb.is_autosome, sb.is_chr_y, sb.start, sb.stop, sb.count, sb.gc)
    k = o["kept_index"]
    # expected text without the product's formatters: the oracle's string-based F2 rounding (oracle/clean.cpp
    # ora_f2_roundtrip: seven significant digits, then half-up on the decimal string) gives the two-decimal value, and
    # Python's own %.2f prints a two-decimal double exactly
    f2 = ora.f2_roundtrip(o["count"])
    exp_text = "".join(f"{sb.names[c]}\t{a}\t{b}\t{'%.2f' % v}\t{g}\n" for c, a, b, v, g in
                       zip(sb.chrom[k].tolist(), sb.start[k].tolist(), sb.stop[k].tolist(), f2.tolist(), sb.gc[k].tolist()))
    assert gzip.open(cleaned, "rt").read() == exp_text
    assert fileio.read_metric(lsd, "localSD") == float(fileio.dotnet_double(o["local_sd"]))

    assert modules.main(["CanvasPartition", "-i", cleaned, "-v", vaf, "-o", part, "-r", str(tmp_path), "-g",
                         f"--evenness-metric-file={evn}", f"--config={cfg}"]) == 0
    order, start, end, cov = fileio.read_cleaned_for_partition(cleaned)
    lens = [len(cov[c]) for c in order]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    p = ora.partition_wavelet(off, np.concatenate([cov[c] for c in order]), is_germline=True, evenness_window=6000,
                              n_threads=4)
    seg = {c: fileio.derive_segments(p["breakpoints"][i], lens[i], start[c], end[c]) for i, c in enumerate(order)}
    # expected .partitioned text, again without the product's writer: double.ToString() of a two-decimal coverage is its
    # shortest decimal form (Segmentation.cs:247), i.e. %.2f with trailing zeros dropped
    segs = fileio.post_process_segments(order, seg, start, end, cov)
    got_rows = [ln.split("\t") for ln in gzip.open(part, "rt").read().splitlines()]
    exp_rows = []
    for c in order:
        for a, b, v in zip(start[c].tolist(), end[c].tolist(), cov[c].tolist()):
            exp_rows.append([c, str(a), str(b), ("%.2f" % v).rstrip("0").rstrip(".")])
    assert [r[:4] for r in got_rows] == exp_rows
    expected = str(tmp_path / "exp.partitioned")
    fileio.write_partitioned(expected, order, segs)
    assert [r[4] for r in got_rows] == [ln.split("\t")[4] for ln in gzip.open(expected, "rt").read().splitlines()]
    got_ev = fileio.read_metric(evn, "evenness")
    assert abs(got_ev - p["evenness"]) <= 1e-9 * abs(p["evenness"])
    ids = [int(l.split("\t")[4]) for l in gzip.open(part, "rt").read().splitlines()]
    assert ids == sorted(ids) and len(set(ids)) > len(order)  # genome-wide running segment counter


def test_partition_without_vaf_gives_one_segment_per_chromosome(tmp_path, engine):
    # WaveletsRunner.cs:74-79: no -v file -> no segments -> every chromosome stays one segment (id -1)
    modules._engine = engine
    s = synth.make_sample(config=2, sample=12, scale=0.01)
    cleaned = str(tmp_path / "s.cleaned")
    part = str(tmp_path / "s.partitioned")
    fileio.write_binned(cleaned, s.names, s.chrom, s.start, s.stop, s.count, s.gc)
    assert modules.main(["CanvasPartition", "-i", cleaned, "-o", part, "-r", str(tmp_path)]) == 0
    ids = {int(l.split("\t")[4]) for l in gzip.open(part, "rt").read().splitlines()}
    assert ids == {-1}


def test_partition_cbs_on_files(tmp_path, engine):
    # CanvasPartition -m CBS on two samples (CanvasPartition.cs:130-149): per-sample CBS, merged boundaries
    modules._engine = engine
    cleaned, parts, per_sample, inputs = [], [], [], []
    for k in range(2):
        s = synth.make_sample(config=2, sample=20 + k, scale=0.004, n_events=4000, chromosomes=["chr1", "chr2", "chrX"])
        path = str(tmp_path / f"s{k}.cleaned")
        fileio.write_binned(path, s.names, s.chrom, s.start, s.stop, s.count.astype(np.float32), s.gc)
        cleaned.append(path)
        parts.append(str(tmp_path / f"s{k}.partitioned"))
        order, start, end, cov = fileio.read_cleaned_for_partition(path)
        lens = [len(cov[c]) for c in order]
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        r = ora.partition_cbs(off, np.concatenate([cov[c] for c in order]))
        per_sample.append({c: [(int(start[c][a]), int(end[c][b])) for a, b in zip(r["segments"][i]["first"], r["segments"][i]["last"])]
                           for i, c in enumerate(order)})
        inputs.append((order, start, end, cov))
    argv = ["CanvasPartition", "-m", "CBS", "-r", str(tmp_path)]
    for c, p in zip(cleaned, parts):
        argv += ["-i", c, "-o", p]
    assert modules.main(argv) == 0
    merged = fileio.split_overlapping_segments(per_sample)
    assert sum(len(v) for v in merged.values()) > max(sum(len(v) for v in ps.values()) for ps in per_sample) > 3
    for (order, start, end, cov), p in zip(inputs, parts):
        expected = str(tmp_path / "exp.partitioned")
        fileio.write_partitioned(expected, order, fileio.post_process_segments(order, merged, start, end, cov))
        assert gzip.open(p, "rt").read() == gzip.open(expected, "rt").read()
