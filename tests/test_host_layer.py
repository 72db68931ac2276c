"""Host-side logic (no GPU): file formats, .NET text semantics, option parsing, segment
post-processing against the reference's own SegmentationResultsProcessorTests vectors."""
import gzip
import os

import numpy as np
import pytest

from canvas_b200 import fileio, modules, multi, synth, textcodec
from oracle import pyoracle as ora


def test_f2_roundtrip_matches_oracle_string_implementation():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(0, 300, 50000), rng.uniform(0, 1, 5000), 10 ** rng.uniform(-6, 7, 10000),
                        np.arange(0, 2000) / 32.0 + 100, np.arange(0, 2000) / 64.0 + 10,
                        [0, 0.005, 0.004, 0.015, 99.995, 9.995, 999999.5, 1e7, -3.14159]]).astype(np.float32)
    assert np.array_equal(ora.f2_roundtrip(x), textcodec.f2_roundtrip(x))
    assert textcodec.f2_text(np.array([95.02041, 0.005, 0.004, 12.345, 1e6], np.float32)) == \
        ["95.02", "0.01", "0.00", "12.35", "1000000.00"]


def test_post_process_segments_reference_vectors():
    # CanvasTest/CanvasPartition/SegmentationResultsProcessorTests.cs:10-95
    order = ["chr1"]
    segs = {"chr1": [(1, 1000), (1100, 4500), (4600, 5000)]}
    start = {"chr1": np.array([100, 600, 1200, 1300, 4001, 5000])}
    end = {"chr1": np.array([500, 890, 1299, 4000, 4500, 5050])}
    cov = {"chr1": np.array([10, 10, 50, 100, 25, 10], float)}

    def summary(res):
        out = []
        for s in res["chr1"]:
            b = s["bins"]
            out.append((min(x[0] for x in b), max(x[1] for x in b), float(np.median([x[2] for x in b])), len(b)))
        return out

    r = fileio.post_process_segments(order, segs, start, end, cov, {}, 100)
    # the reference test asserts 3 segments with the first one spanning two bins although no bin starts
    # at a segment start: bins before the first forced split share segment id -1 ... but MaxInterBinDist
    # (100) splits where the gap exceeds it
    assert summary(r) == [(100, 890, 10.0, 2), (1200, 4500, 50.0, 3), (5000, 5050, 10.0, 1)]
    r = fileio.post_process_segments(order, segs, start, end, cov, {"chr1": [(525, 575)]}, 100)
    assert summary(r) == [(100, 500, 10.0, 1), (600, 890, 10.0, 1), (1200, 4500, 50.0, 3), (5000, 5050, 10.0, 1)]
    r = fileio.post_process_segments(order, segs, start, end, cov, {"chr1": [(585, 635)]}, 100)
    assert summary(r) == [(100, 500, 10.0, 1), (600, 890, 10.0, 1), (1200, 4500, 50.0, 3), (5000, 5050, 10.0, 1)]


def test_derive_segments():
    start = np.arange(0, 20000, 1000)
    end = start + 1000
    assert fileio.derive_segments([0, 5, 12], 20, start, end) == [(0, 5000), (5000, 12000), (12000, 20000)]
    assert fileio.derive_segments([0], 20, start, end) == [(0, 20000)]          # < 2 breakpoints
    assert fileio.derive_segments([0, 3], 10, start[:10], end[:10]) == [(0, 10000)]  # <= 10 bins


def test_bin_filter_semantics():
    f = fileio.BinFilter({"chr1": [(100, 200), (500, 600)]})
    assert [f.skip("chr1", a, b) for a, b in [(0, 100), (50, 101), (199, 250), (200, 300), (599, 700), (600, 700)]] == \
        [False, True, True, False, True, False]
    assert f.skip("chr2", 0, 1000) is False
    assert f.skip("chr1", 150, 160) is True  # chromosome switch resets the cursor


def test_binned_roundtrip_and_run_ids(tmp_path):
    p = str(tmp_path / "x.binned")
    names = ["chr1", "chrX"]
    chrom = np.array([0, 0, 1, 1], np.uint8)
    start = np.array([0, 1000, 0, 1000], np.int32)
    count = np.array([95.02041, 100.125, 0.0, 7.005], np.float32)
    fileio.write_binned(p, names, chrom, start, start + 1000, count, np.array([40, 41, 42, 43], np.uint8))
    lines = gzip.open(p, "rt").read().splitlines()
    assert lines[0] == "chr1\t0\t1000\t95.02\t40" and lines[3].split("\t")[3] in ("7.00", "7.01")
    s = fileio.read_binned(p)
    assert s.names == names and s.chrom.tolist() == [0, 0, 1, 1] and s.is_autosome.tolist() == [1, 0]
    assert np.array_equal(s.count.astype(np.float64), textcodec.f2_roundtrip(count).astype(np.float32).astype(np.float64))


def test_dotnet_double_and_metric_files(tmp_path):
    assert fileio.dotnet_double(96.32670330162374) == "96.3267033016237"
    assert fileio.dotnet_double(95.02) == "95.02" and fileio.dotnet_double(100.0) == "100"
    p = str(tmp_path / "m.txt")
    fileio.write_metric(p, "localSD", 1.8605272624936706)
    assert open(p).read() == "#localSD\t1.86052726249367\n"
    assert abs(fileio.read_metric(p, "localSD") - 1.86052726249367) < 1e-15


def test_option_parsing_like_ndesk():
    o = modules._parse(["-i", "a", "-o=b", "-g", "-s", "--local-sd-metric-file=f", "-w", "50"], modules.CLEAN_SPEC)
    assert o == {"infile": "a", "outfile": "b", "gcnorm": True, "filtsize": True, "local_sd_file": "f", "weightedmedian": "50"}
    with pytest.raises(modules.UnknownArguments):
        modules._parse(["-i", "a", "stray"], modules.CLEAN_SPEC)
    o = modules._parse(["-i", "a", "-i", "b", "-o", "c", "-r", "ref", "-g"], modules.PARTITION_SPEC)
    assert o["infile"] == ["a", "b"] and o["germline"] is True


def test_module_exit_codes_without_gpu(tmp_path, capsys):
    assert modules.main(["CanvasClean"]) == 0                       # help when -i/-o are missing (:461-465)
    assert "Usage: CanvasClean.exe" in capsys.readouterr().out
    assert modules.main(["CanvasClean", "-i", str(tmp_path / "nope"), "-o", "x"]) == 1   # :468-472
    assert modules.main(["CanvasPartition", "-i", str(tmp_path / "nope"), "-o", "x", "-r", "ref"]) == 1
    assert modules.main(["CanvasClean", "-i", "a", "-o", "b", "bogus"]) == 255


def test_lpt_assignment_balances():
    lengths = [int(l / 1000) for _, l in synth.HG19]
    owner = multi.assign_chromosomes_lpt(lengths, 8)
    load = np.bincount(owner, weights=lengths, minlength=8)
    assert load.max() / sum(lengths) < 0.14   # 8-way makespan close to 1/8 (~12.6 %)
    assert sorted(set(owner.tolist())) == list(range(8))


def test_pack_unpack_breakpoints():
    bps = [np.array([0, 5, 9], np.int32), np.zeros(0, np.int32), np.array([0], np.int32)]
    buf = multi.pack_breakpoints(bps, 64)
    assert buf[0] == 4
    out = multi.unpack_breakpoints(buf[None, :], 3)
    assert [b.tolist() for b in out] == [[0, 5, 9], [], [0]]


def test_split_overlapping_segments_reference_vectors():
    # CanvasTest/CanvasPartition/GenomeSegmentationResultsTests.cs
    f = fileio.split_overlapping_segments
    one = {"chr1": [(1, 200)]}
    assert f([one]) == one
    assert f([one, {"chr1": [(1, 200)]}]) == one
    assert f([{"chr1": [(1, 300)]}, {"chr1": [(1, 200), (200, 300)]}]) == {"chr1": [(1, 200), (200, 300)]}
    assert f([{"chr1": [(0, 200)]}, {"chr1": [(100, 300)]}]) == {"chr1": [(0, 100), (100, 200), (200, 300)]}
    assert f([{"chr1": [(0, 100)]}, {"chr1": [(0, 200)]}]) == {"chr1": [(0, 100), (100, 200)]}
    assert f([{"chr1": [(0, 300)]}, {"chr1": [(100, 200)]}]) == {"chr1": [(0, 100), (100, 200), (200, 300)]}
    assert f([{"chr1": [(0, 600)]}, {"chr1": [(100, 200), (200, 300), (500, 700), (800, 900)]}]) == {
        "chr1": [(0, 100), (100, 200), (200, 300), (300, 500), (500, 600), (600, 700), (800, 900)]}
    assert f([{"chr1": [(0, 100)]}, {"chr1": [(100, 200)]}]) == {"chr1": [(0, 100), (100, 200)]}
    assert f([{"chr1": [(0, 100)], "chr2": [(300, 400)]}, {"chr1": [(100, 200)], "chr2": [(500, 600)]}]) == {
        "chr1": [(0, 100), (100, 200)], "chr2": [(300, 400), (500, 600)]}


def test_cbs_segments_from_lengths():
    start = np.array([0, 10, 20, 30, 40]); end = start + 10
    assert fileio.cbs_segments([2, 3], start, end) == [(0, 20), (20, 50)]


def test_float_default_text_and_roundtrip():
    """float.ToString() of .NET Core 2.0 (7 significant digits) as used by NormalizeCanvasClean."""
    from canvas_b200 import textcodec
    v = np.array([0.0, 1.0, 99.46, 123.45, 12345.68, 123456.78, 1234567.0, 12345678.0, 0.0001, 0.00001234, 100.0], np.float32)
    txt = textcodec.float_default_text(v)
    assert txt == ["0", "1", "99.46", "123.45", "12345.68", "123456.8", "1234567", "1.234568E+07", "0.0001", "1.234E-05", "100"]
    rt = textcodec.float_default_roundtrip(v)
    assert rt.tolist() == [float(t) for t in txt]
    rng = np.random.default_rng(0)
    w = np.round(rng.gamma(2, 3000, 20000), 2).astype(np.float32)
    assert textcodec.float_default_roundtrip(w).tolist() == [float(t) for t in textcodec.float_default_text(w)]


def test_merge_oracle_follows_dictionary_semantics():
    from oracle import pyoracle as ora
    a = [("chr1", 0, 10, 1.0), ("chr1", 10, 20, 2.0), ("chr2", 0, 10, 3.0)]
    b = [("chr1", 10, 25, 5.0), ("chr2", 0, 10, 6.0), ("chr3", 0, 10, 7.0)]
    out = ora.merge_multi_sample_cleaned([a, b])
    assert [(c, s, e) for c, s, e, _ in out] == [("chr1", 10, 25), ("chr2", 0, 10)]  # the last file's stop wins
    assert [[float(x) for x in v] for *_, v in out] == [[2.0, 5.0], [3.0, 6.0]]


def test_native_text_codec_matches_python_codec(tmp_path):
    """cg_format_bins / cg_parse_bins (host code of the library) against the Python restatement of .NET's formats."""
    from canvas_b200 import fileio, native, textcodec
    rng = np.random.default_rng(3)
    n = 60000
    names = ["chr1", "chr2", "chrX", "chr1"]  # chr1 reappears: a new run
    chrom = np.sort(rng.integers(0, 4, n)).astype(np.uint8)
    start = (np.arange(n) * 1000).astype(np.int32)
    stop = start + 1000
    count = np.concatenate([rng.gamma(2, 60, n - 8), [0.0, 0.005, 0.015, 1e7, 123456.785, 2.5e-5, 99.995, 1234567.9]]).astype(np.float32)
    gc = rng.integers(0, 101, n).astype(np.uint8)
    txt = native.format_bins(names, chrom, start, stop, count, gc, n_threads=3)
    lines = txt.decode().splitlines()
    assert len(lines) == n
    want = textcodec.f2_text(count)
    for i in list(range(0, n, 997)) + list(range(n - 8, n)):
        assert lines[i] == f"{names[chrom[i]]}\t{start[i]}\t{stop[i]}\t{want[i]}\t{gc[i]}"
    t4 = native.format_bins(names, chrom, start, stop, count, None, four_columns=True).decode().splitlines()
    g7 = textcodec.float_default_text(count)
    for i in list(range(0, n, 997)) + list(range(n - 8, n)):
        assert t4[i].split("\t")[3] == g7[i]
    rn, rc, ra, rb, rv, rg = native.parse_bins(txt, n_threads=4)
    assert rn == ["chr1", "chr2", "chrX", "chr1"][:len(rn)] and len(rn) == len(np.unique(chrom))
    assert np.array_equal(ra, start) and np.array_equal(rb, stop) and np.array_equal(rg, gc)
    assert np.array_equal(rv, np.array([float(x) for x in want]).astype(np.float32))
    assert np.array_equal(np.diff(rc.astype(int)) != 0, np.diff(chrom.astype(int)) != 0)
    # through the file layer (gzip) and with CRLF / blank lines / four columns
    p = tmp_path / "x.binned"
    fileio.write_binned(str(p), names, chrom, start, stop, count, gc)
    s = fileio.read_binned(str(p))
    assert np.array_equal(s.start, start) and np.array_equal(s.gc, gc) and len(s.names) == len(rn)
    rn2, rc2, ra2, rb2, rv2, rg2 = native.parse_bins(b"chr1\t0\t10\t1.50\r\n\nchr2\t10\t20\t2.25\t7\n")
    assert rn2 == ["chr1", "chr2"] and ra2.tolist() == [0, 10] and rv2.tolist() == [1.5, 2.25] and rg2.tolist() == [0, 7]
    with pytest.raises(ValueError):
        native.parse_bins(b"chr1\t0\t10\n")
    with pytest.raises(ValueError):
        native.parse_bins(b"chr1\t0\tx\t1.0\t3\n")


def test_ploidy_info_and_ploidy_splits(tmp_path):
    # PloidyInfo.cs:56-109 and the ploidy rule of SegmentationResultsProcessor.IsNewSegment (:116-125)
    vcf = tmp_path / "ploidy.vcf"
    vcf.write_text("##fileformat=VCFv4.1\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\n"
                   "chrX\t1\t.\tN\t<CNV>\t.\tPASS\tEND=2500\tCN\t1\n"
                   "chrX\t2501\t.\tN\t<CNV>\t.\tPASS\tEND=9000\tCN\t.\n"
                   "chrY\t1\t.\tN\t<CNV>\t.\tPASS\tEND=100000\tGT:CN\t0:0\n")
    p = fileio.PloidyInfo.load_vcf_no_sample_id(str(vcf))
    assert p.by_chr == {"chrX": [(1, 2500, 1), (2501, 9000, 2)], "chrY": [(1, 100000, 0)]}
    assert p.reference_copy_number("chr1", 0, 1000) == 2          # chromosome not listed
    assert p.reference_copy_number("chrX", 0, 1000) == 1
    assert p.reference_copy_number("chrX", 2000, 3000) == 1        # 500 bases each: the first best count wins
    assert p.reference_copy_number("chrX", 2001, 3000) == 2        # 499 against 500
    assert p.reference_copy_number("chrY", 500, 600) == 0
    assert p.is_uniform("chr1", 1, 10 ** 6) and p.is_uniform("chrX", 1, 2500) and not p.is_uniform("chrX", 2500, 2501)
    # an empty PloidyInfo (what the reference's tests pass) never splits
    order = ["chrX"]
    start = {"chrX": np.array([0, 1000, 2000, 3000, 4000])}
    end = {"chrX": np.array([1000, 2000, 3000, 4000, 5000])}
    cov = {"chrX": np.arange(5.0)}
    ids = lambda r: [[b[0] for b in s["bins"]] for s in r["chrX"]]
    assert ids(fileio.post_process_segments(order, {}, start, end, cov, {}, 10 ** 6, fileio.PloidyInfo())) == [[0, 1000, 2000, 3000, 4000]]
    # the bin [2000, 3000) spans the change at 2500/2501: it opens a new segment; the next bin is checked from the
    # previous bin's end (3000) and is uniform again
    assert ids(fileio.post_process_segments(order, {}, start, end, cov, {}, 10 ** 6, p)) == [[0, 1000], [2000, 3000, 4000]]
    # two genotype columns are refused when no sample id is given
    bad = tmp_path / "two.vcf"
    bad.write_text("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\tS2\n")
    with pytest.raises(ValueError):
        fileio.PloidyInfo.load_vcf_no_sample_id(str(bad))


def test_on_target_mask_follows_the_reference_walk():
    # BinCounts.LoadBinCounts (BinCounts.cs:102-160): 1-based regions, 0-based half-open bins
    names = ["chr1", "chr2", "chr3"]
    chrom = [0, 0, 0, 0, 1, 1, 2, 0]
    start = [0, 100, 200, 300, 0, 100, 0, 400]
    stop = [100, 200, 300, 400, 100, 200, 100, 500]
    regions = {"chr1": [(150, 180), (301, 310), (450, 460)], "chr3": [(101, 200)]}
    m = fileio.on_target_mask(names, chrom, start, stop, regions)
    # chr1: [100,200) holds 150-180; [200,300): the next region starts at 301 > 300 (off); [300,400) holds 301;
    # chr2 has no regions; chr3's region starts after the bin's last base; the chromosome reappearing restarts its cursor
    assert m.tolist() == [0, 1, 0, 1, 0, 0, 0, 1]
    # a region ending exactly on the bin's first base (End == Start + 1) still overlaps; one ending before does not
    assert fileio.on_target_mask(["c"], [0, 0], [10, 20], [20, 30], {"c": [(5, 11), (12, 20)]}).tolist() == [1, 0]


def test_canvas_normalize_argument_handling(tmp_path, capsys):
    # the checks of CanvasNormalizeParameters.ParseCommandLine (CanvasNormalize/Program.cs:45-148) run before any device work
    t = tmp_path / "t.binned"
    t.write_text("chr1\t0\t100\t5\t40\n")
    assert modules.canvas_normalize_main(["-n", str(t), "-o", str(tmp_path / "o")]) == 1            # no tumour file
    assert modules.canvas_normalize_main(["-t", str(t), "-o", str(tmp_path / "o")]) == 1            # no normal file
    assert modules.canvas_normalize_main(["-t", str(t), "-n", str(t)]) == 1                         # no output
    assert modules.canvas_normalize_main(["-t", str(t), "-n", str(t), "-o", "o", "-r", "5"]) == 1   # -r once
    assert modules.canvas_normalize_main(["-t", str(tmp_path / "nope"), "-n", str(t), "-o", "o"]) == 1
    assert modules.canvas_normalize_main(["-t", str(t), "-n", str(t), "-n", str(t), "-o", "o", "-m", "PCA"]) == 1
    with pytest.raises(modules.UnknownArguments):
        modules.canvas_normalize_main(["-t", str(t), "--bogus"])
    with pytest.raises(ValueError):
        modules.canvas_normalize_main(["-t", str(t), "-n", str(t), "-o", "o", "-m", "Nonsense"])
    out = capsys.readouterr()
    assert "Please specify the tumor bed file." in out.err and "does not exist! Exiting." in out.out


def test_ploidy_counts_on_the_reference_ploidy_test_scenarios():
    # The interval scenarios of CanvasTest/CanvasCommon/ReferencePloidyTests.cs (they exercise the sibling class
    # ReferencePloidy; PloidyInfo.getPloidyCounts, PloidyInfo.cs:92-109, must split the same one-based intervals the same way)
    def info(*intervals):
        p = fileio.PloidyInfo()
        p.by_chr["chrX"] = list(intervals)
        return p
    assert fileio.PloidyInfo().is_uniform("chrX", 1, 2)                                   # EmptyVcf_ReferencePloidyIs2
    assert info((1, 2, 1))._counts("chrX", 1, 2) == [0, 2, 0, 0, 0]                        # VcfPloidy1AndSameQueryInterval
    assert info((1, 1, 1))._counts("chrX", 1, 2) == [0, 1, 1, 0, 0]                        # PartialOverlap: ploidy 1 and 2
    assert not info((1, 1, 1)).is_uniform("chrX", 1, 2)
    assert info((1, 1, 1), (2, 2, 1)).is_uniform("chrX", 1, 2)                             # two adjacent ploidy-1 intervals
    assert info((1, 1, 1), (2, 2, 1), (3, 3, 1), (4, 4, 1))._counts("chrX", 1, 4) == [0, 4, 0, 0, 0]
    assert info((2, 2, 1), (4, 4, 3))._counts("chrX", 1, 5) == [0, 1, 3, 1, 0]             # MultiplePloidyAndLargeQuery
    assert info((1, 4, 1)).reference_copy_number("chrX", 1, 3) == 1                        # query (2, 3) inside a ploidy-1 region
    assert info((1, 4, 1)).is_uniform("chrX", 2, 3)


def test_parallel_gzip_is_one_standard_member():
    import subprocess
    import zlib
    rng = np.random.default_rng(0)
    lines = [f"chr{1 + i % 22}\t{i * 1000}\t{i * 1000 + 1000}\t{rng.integers(0, 300)}.{rng.integers(0, 100):02d}\t{rng.integers(20, 70)}\n"
             for i in range(60000)]
    text = "".join(lines).encode()
    for chunk, threads in ((1 << 16, 4), (100000, 3), (1 << 30, 4), (1 << 16, 1)):
        z = fileio.gzip_bytes(text, chunk=chunk, threads=threads)
        assert gzip.decompress(z) == text
        d = zlib.decompressobj(31)                 # a single gzip member: nothing is left over after it
        assert d.decompress(z) == text and d.eof and d.unused_data == b""
        assert z[:4] == b"\x1f\x8b\x08\x00"
    assert gzip.decompress(fileio.gzip_bytes(b"")) == b""
    # the system gzip agrees (when there is one)
    try:
        p = subprocess.run(["gzip", "-t"], input=fileio.gzip_bytes(text, chunk=1 << 16, threads=4), capture_output=True)
        assert p.returncode == 0, p.stderr
    except FileNotFoundError:
        pass


def test_fast_cleaned_reader_equals_the_line_loop(tmp_path):
    rng = np.random.default_rng(5)
    n = 20000
    chrom = np.sort(rng.integers(0, 5, n))
    names = ["chr1", "chr2", "chrX", "chrY", "chrM"]
    start = np.arange(n) * 1000
    cov_text = [f"{rng.integers(0, 400)}.{rng.integers(0, 100):02d}" for _ in range(n)]
    cov_text[5], cov_text[6], cov_text[7], cov_text[8] = "1E-05", "1.234568E+07", "0", "123456.78"
    five = "".join(f"{names[c]}\t{a}\t{a + 1000}\t{t}\t{40 + a % 7}\n" for c, a, t in zip(chrom, start, cov_text))
    four = "".join(f"{names[c]}\t{a}\t{a + 1000}\t{t}\n" for c, a, t in zip(chrom, start, cov_text))
    # chromosome that reappears later in the file, CRLF line ends, a blank line
    tricky = five + "chr1\t0\t10\t5.5\t40\r\n\nchr2\t20\t30\t6.25\t41\n"
    for k, (body, gz) in enumerate(((five, True), (four, False), (tricky, True))):
        p = tmp_path / f"f{k}"
        p.write_bytes(gzip.compress(body.encode()) if gz else body.encode())
        want = fileio._read_cleaned_rows(str(p))
        got = fileio.read_cleaned_for_partition(str(p))
        assert want[0] == got[0]
        for w, g in zip(want[1:], got[1:]):
            assert list(w) == list(g)
            for c in w:
                assert w[c].dtype == g[c].dtype and np.array_equal(w[c], g[c]), c
    # padded fields and a -b filter take the line loop; a malformed line raises as before
    p = tmp_path / "padded"
    p.write_text(" chr1 \t0\t10\t5.5\t40\n")
    assert fileio.read_cleaned_for_partition(str(p))[0] == ["chr1"]
    p.write_text("chr1\tx\t10\t5.5\t40\n")
    with pytest.raises(ValueError):
        fileio.read_cleaned_for_partition(str(p))


def test_write_partitioned_text(tmp_path):
    order = ["chr1", "chr2"]
    segs = {"chr1": [{"id": 0, "bins": [(1000, 2000, 95.5), (0, 1000, 95.02)]}, {"id": 1, "bins": [(2000, 3000, 1e-05)]}],
            "chr2": [{"id": 2, "bins": [(0, 1000, 100.0), (1000, 2000, 95.02), (2000, 3000, float("nan"))]}]}
    p = tmp_path / "x.partitioned"
    fileio.write_partitioned(str(p), order, segs)
    assert gzip.open(p, "rt").read() == ("chr1\t0\t1000\t95.02\t0\nchr1\t1000\t2000\t95.5\t0\nchr1\t2000\t3000\t1E-05\t1\n"
                                         "chr2\t0\t1000\t100\t2\nchr2\t1000\t2000\t95.02\t2\nchr2\t2000\t3000\tNaN\t2\n")


@pytest.mark.parametrize("extra,samples", [([], 1), (["--gpus", "2"], 2), (["--config", "3"], 2), (["--config", "4"], 3)])
def test_bench_reference_arm_contract(extra, samples):
    # `bench.py --impl reference` needs no GPU: one JSON line with the contract's keys (a shrunken genome keeps it to seconds);
    # it processes as many samples as the GPU arm of the same command line (one per GPU, the pair, the trio)
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--scale", "0.05"] + extra, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mbins/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d, k
    assert "workload" in d["config"] and d["config"]["samples"] == samples


def test_numpy_merge_equals_the_dictionary_merge():
    # the CPU arm of config 4 uses the column-array merge; it must select what Utilities.MergeMultiSampleCleanedBedFile selects
    from oracle import pyoracle as ora
    rng = np.random.default_rng(8)
    samples = []
    for k in range(3):
        keep = rng.random(600) > 0.15
        chrom = np.repeat(np.arange(3), 200)[keep].astype(np.uint8)
        start = (np.tile(np.arange(200), 3) * 1000)[keep].astype(np.int32)
        samples.append((chrom, start, (start + 1000 + k).astype(np.int32), rng.normal(100, 10, keep.sum()).astype(np.float32)))
    want = ora.merge_multi_sample_cleaned([[(int(c), int(a), int(b), float(v)) for c, a, b, v in zip(*s)] for s in samples])
    got = ora.merge_common_bins_np(samples)
    assert len(want) == len(got["kept_index"]) > 100
    c0, a0 = samples[0][0][got["kept_index"]], samples[0][1][got["kept_index"]]
    assert [(w[0], w[1], w[2]) for w in want] == list(zip(c0.tolist(), a0.tolist(), got["stop"].tolist()))
    assert np.array_equal(np.array([w[3] for w in want], np.float32).T, got["count"])


def test_chromosome_order_merged_over_pedigree_files():
    # ids by first appearance over all files break the (id, start) order when the first file lost a whole chromosome that a
    # later file still has mid-genome; the merged order keeps every file ascending
    m = fileio.merge_chromosome_orders([["chr1", "chr2", "chrX"], ["chr1", "chr2", "chr3", "chrX", "chrY"], ["chr2", "chr3", "chrY"]])
    assert m == ["chr1", "chr2", "chr3", "chrX", "chrY"]
    for order in (["chr1", "chr2", "chrX"], ["chr1", "chr2", "chr3", "chrX", "chrY"], ["chr2", "chr3", "chrY"]):
        idx = [m.index(c) for c in order]
        assert idx == sorted(idx)
    assert fileio.merge_chromosome_orders([[], ["a"], ["b", "a"]]) == ["b", "a"]


def test_more_than_256_contigs_is_reported_not_garbled(tmp_path):
    # the chromosome cap (8-bit ids) is a documented limit: the readers say so instead of failing somewhere downstream
    lines = "".join(f"contig{c}\t0\t1000\t100.00\t40\n" for c in range(300))
    p = tmp_path / "many.binned"
    p.write_bytes(fileio.gzip_bytes(lines.encode()))
    with pytest.raises(ValueError, match="256 chromosome"):
        fileio.read_binned(str(p))


def test_bind_host_near_gpu_is_only_a_hint():
    # NVML may be missing or report nothing (no GPU here): the call must not raise and must leave the affinity usable
    import os
    from canvas_b200 import native
    before = os.sched_getaffinity(0)
    cpus = native.bind_host_near_gpu(0)
    after = os.sched_getaffinity(0)
    assert cpus is None or (set(cpus) == after and after <= before)
    os.sched_setaffinity(0, before)


def test_segment_with_bins_reference_vector():
    # SegmentWithBinsTests.AddBinTest (CanvasTest/CanvasPartition/SegmentWithBinsTests.cs:22-46): extent and median coverage
    # after every AddBin, and independence of the insertion order
    from canvas_b200 import fileio
    b1, b2, b3 = (100, 2000, 10.0), (2500, 3000, 5.0), (5000, 8000, 45.0)
    seg = {"id": 1, "bins": [b1]}
    assert fileio.segment_extent(seg) == (100, 2000) and fileio.segment_median_coverage(seg) == 10 and len(seg["bins"]) == 1
    seg["bins"].append(b2)
    assert fileio.segment_extent(seg) == (100, 3000) and fileio.segment_median_coverage(seg) == 7.5
    seg["bins"].append(b3)
    assert fileio.segment_extent(seg) == (100, 8000) and fileio.segment_median_coverage(seg) == 10 and len(seg["bins"]) == 3
    other = {"id": 1, "bins": [b1, b3, b2]}
    assert fileio.segment_extent(other) == (100, 8000) and fileio.segment_median_coverage(other) == 10


def test_mode_parsers_reference_vectors():
    # TestUtilities.TestParseCanvasNormalizeMode(+_WithBadString) (CanvasTest/TestUtilities.cs:12-31) and the sibling parser of
    # CanvasClean's -m (CanvasCommon/Utilities.cs:91-102): case-insensitive, trimmed, unknown names throw
    import pytest
    from canvas_b200 import modules
    for text, want in (("weightedaverage", "weightedaverage"), ("WeightedAverage", "weightedaverage"), ("bestlr2", "bestlr2"),
                       ("BestLR2", "bestlr2"), ("pca", "pca"), ("PCA", "pca"), (" PCA ", "pca")):
        assert modules.parse_canvas_normalize_mode(text) == want
    with pytest.raises(ValueError):
        modules.parse_canvas_normalize_mode("badmode")
    assert modules.parse_gc_normalization_mode("MedianByGC") == "medianbygc" and modules.parse_gc_normalization_mode("LOESS ") == "loess"
    with pytest.raises(ValueError):
        modules.parse_gc_normalization_mode("median")
