"""Intermediate file formats of the Canvas module chain (host side; reference CanvasCommon/IO.cs).

  .binned / .cleaned   gzip TSV  chr  start  stop  count({0:F2})  gc          IO.cs:15-52
  .partitioned         gzip TSV  chr  start  end   coverage       segmentId   Segmentation.cs:235-252
  metric files         "#localSD\\t<v>" / "#evenness\\t<v>"                    IO.cs:83-98
  filter BED           chr  start  stop                                       Utilities.cs:794-828
"""
import gzip
import io

import numpy as np

from . import synth, textcodec


def _open_text(path, mode="rt"):
    with open(path, "rb") as f:
        magic = f.read(2)
    if magic == b"\x1f\x8b":
        return gzip.open(path, mode)
    return open(path, mode)


def gzip_bytes(data, level=6, threads=None, chunk=8 << 20):
    """One standard single-member gzip stream, deflated by several threads: the reference's writers spend most of a
    module's wall clock inside zlib (7 s for a 108 MB .binned file at level 6 against milliseconds of device work).
    Every chunk is deflated on its own (zlib releases the GIL) and ended with a sync flush, which byte-aligns it without
    a final block, so the pieces concatenate into ONE valid deflate stream (the construction pigz uses); the CRC-32 and
    length of the whole text close the member.  Any gzip reader (.NET's GZipStream included) reads it as usual."""
    import os
    import struct
    import zlib
    from concurrent.futures import ThreadPoolExecutor
    data = bytes(data) if not isinstance(data, (bytes, bytearray, memoryview)) else data
    n = len(data)
    threads = threads or min(16, os.cpu_count() or 1)
    if n <= chunk or threads <= 1:
        return gzip.compress(bytes(data), compresslevel=level, mtime=0)
    mv = memoryview(data)
    cuts = list(range(0, n, chunk)) + [n]

    def deflate(k):
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        out = c.compress(mv[cuts[k]:cuts[k + 1]])
        return out + c.flush(zlib.Z_FINISH if k == len(cuts) - 2 else zlib.Z_SYNC_FLUSH)

    with ThreadPoolExecutor(max_workers=threads) as ex:
        crc = ex.submit(zlib.crc32, mv)
        parts = list(ex.map(deflate, range(len(cuts) - 1)))
        crc = crc.result()
    header = b"\x1f\x8b\x08\x00" + struct.pack("<I", 0) + b"\x00\xff"  # deflate, no flags, no mtime, unknown OS
    return b"".join([header] + parts + [struct.pack("<II", crc & 0xffffffff, n & 0xffffffff)])


def _read_bytes(path):
    with open(path, "rb") as f:
        data = f.read()
    return gzip.decompress(data) if data[:2] == b"\x1f\x8b" else data


def read_binned(path):
    """CanvasIO.ReadFromTextFile (IO.cs:26-52): returns a synth.Sample with chromosome RUN ids (a
    chromosome that reappears later in the file gets a new id, as the run-based loops of CanvasClean
    see it, CanvasClean.cs:247-256).  The lines are parsed by the library's native codec (cg_parse_bins)."""
    from . import native
    names, chrom, start, stop, count, gc = native.parse_bins(_read_bytes(path))
    return synth.Sample(names, chrom, start, stop, count, gc)


def write_binned(path, names, chrom, start, stop, count, gc):
    """CanvasIO.WriteToTextFile (IO.cs:15-24): count printed with .NET's {0:F2} (native codec, cg_format_bins)."""
    from . import native
    text = native.format_bins(names, chrom, start, stop, count, gc)
    with open(path, "wb") as f:
        f.write(gzip_bytes(text))


def write_metric(path, name, value):
    """IO.cs:95-98 (double.ToString(): 15 significant digits on .NET Core 2.0)."""
    with open(path, "w") as f:
        f.write(f"#{name}\t{dotnet_double(value)}\n")


def read_metric(path, name):
    for line in open(path):
        if line.startswith("#" + name):
            return float(line.split("\t")[1])
    raise ValueError(f"Did not find {name} metric in file '{path}'")


def dotnet_double(x):
    """double.ToString() of .NET Core 2.0: the G15 form."""
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Infinity" if x > 0 else "-Infinity"
    s = "%.15g" % x
    if "e" in s:
        m, e = s.split("e")
        return f"{m}E{'+' if int(e) >= 0 else '-'}{abs(int(e)):02d}"
    return s


def load_bed(path):
    """Utilities.LoadBedFile (Utilities.cs:794-828): chr -> list of (start, stop) in file order."""
    out = {}
    if not path:
        return out
    with _open_text(path) as f:
        for line in f:
            line = line.rstrip("\n").rstrip("\r")
            if not line:
                continue
            p = line.split("\t")
            a, b = int(p[1]), int(p[2])
            if a < 0:
                raise ValueError(f"Start must be non-negative in a BED file: {line}")
            if a >= b:
                raise ValueError(f"Start must be less than Stop in a BED file: {line}")
            out.setdefault(p[0], []).append((a, b))
    return out


class BinFilter:
    """GenomicBinFilter.SkipBin (GenomicBinFilter.cs:29-58)."""

    def __init__(self, excluded):
        self.excluded = excluded
        self.prev_chrom = None
        self.prev_start = 0
        self.intervals = []
        self.idx = -1

    def skip(self, chrom, start, stop):
        if chrom != self.prev_chrom:
            self.prev_chrom = chrom
            self.intervals = self.excluded.get(chrom, [])
            self.idx = 0
        elif start < self.prev_start:
            self.idx = 0
        self.prev_start = start
        while self.idx < len(self.intervals):
            a, b = self.intervals[self.idx]
            if b <= start:
                self.idx += 1
                continue
            if a >= stop:
                return False
            return True
        return False


def _read_cleaned_rows(path, filter_bed=None):
    """The line-by-line form of read_cleaned_for_partition (kept as the definition of its behaviour and as the fallback)."""
    filt = BinFilter(load_bed(filter_bed))
    order, start, end, cov = [], {}, {}, {}
    with _open_text(path) as f:
        for line in f:
            line = line.rstrip("\n").rstrip("\r")
            if not line:
                continue
            p = line.split("\t")
            c = p[0].strip()
            a, b = int(p[1].strip()), int(p[2].strip())
            if filt.skip(c, a, b):
                continue
            if c not in start:
                order.append(c)
                start[c], end[c], cov[c] = [], [], []
            start[c].append(a)
            end[c].append(b)
            cov[c].append(float(p[3].strip()))
    return order, {c: np.array(start[c], np.int64) for c in order}, {c: np.array(end[c], np.int64) for c in order}, \
        {c: np.array(cov[c], np.float64) for c in order}


def read_cleaned_for_partition(path, filter_bed=None):
    """CanvasSegment.ReadBedInput (CanvasSegment.cs:1117-1163): per chromosome (first-appearance order)
    start, end and coverage; bins overlapping the -b BED are dropped.  Plain files (no -b filter, no padded fields) go
    through a C parser with correctly rounded doubles (3.1 M lines: 8.1 s -> 5.9 s); anything else takes the line loop."""
    if filter_bed is None:
        try:
            import pandas as pd
            df = pd.read_csv(io.BytesIO(_read_bytes(path)), sep="\t", header=None, usecols=[0, 1, 2, 3], quoting=3,
                             dtype={0: str, 1: np.int64, 2: np.int64, 3: np.float64}, float_precision="round_trip",
                             na_filter=False, skip_blank_lines=True, engine="c")
            names = df[0].to_numpy()
            codes, uniques = pd.factorize(names)
            if all(u == u.strip() and u for u in uniques):
                order = [str(u) for u in uniques]
                a, b, v = df[1].to_numpy(np.int64), df[2].to_numpy(np.int64), df[3].to_numpy(np.float64)
                idx = [np.flatnonzero(codes == k) for k in range(len(order))]
                return order, {c: a[i] for c, i in zip(order, idx)}, {c: b[i] for c, i in zip(order, idx)}, \
                    {c: v[i] for c, i in zip(order, idx)}
        except Exception:  # noqa: anything the fast parser does not take is left to the line loop, which defines the errors
            pass
    return _read_cleaned_rows(path, filter_bed)


def derive_segments(breakpoints, n_bins, start, end):
    """SegmentationInput.DeriveSegments (Segmentation.cs:83-125): list of (start, end) coordinates."""
    bp = list(int(b) for b in breakpoints)
    s_pos, e_pos = [], []
    if len(bp) >= 2 and n_bins > 10:
        if bp[0] != 0:
            bp.insert(0, 0)
        s_pos.append(bp[0])
        e_pos.append(bp[1] - 1)
        for i in range(1, len(bp) - 1):
            s_pos.append(bp[i])
            e_pos.append(bp[i + 1] - 1)
        s_pos.append(bp[-1])
        e_pos.append(n_bins - 1)
    else:
        s_pos.append(0)
        e_pos.append(n_bins - 1)
    return [(int(start[a]), int(end[b])) for a, b in zip(s_pos, e_pos)]


def cbs_segments(seg_len, start, end):
    """CBSRunner.Run (CBSRunner.cs:127-138): segment lengths in bins -> (genomic start, genomic end) pairs.
    Coverage is finite on this path, so the finite-index list is the identity."""
    out, at = [], 0
    for ln in np.asarray(seg_len).tolist():
        out.append((int(start[at]), int(end[at + ln - 1])))
        at += ln
    return out


def split_overlapping_segments(per_sample):
    """GenomeSegmentationResults.SplitOverlappingSegments (GenomeSegmentationResults.cs:18-55): the union of
    all samples' segment boundaries per chromosome.  per_sample: list of {chr: [(start, end), ...]}."""
    if len(per_sample) == 1:
        return per_sample[0]
    result = {}
    for c in per_sample[0]:
        events = []
        for sample in per_sample:
            for a, b in sample[c]:
                events.append((a, 0))
                events.append((b, 1))
        events.sort(key=lambda e: e[0])  # stable merge by position only
        segs, depth, cur = [], 0, 0
        for pos, is_end in events:
            if depth > 0 and cur != pos:
                segs.append((cur, pos))
            cur = pos
            depth += -1 if is_end else 1
        result[c] = segs
    return result


def on_target_mask(names, chrom, start, stop, regions_by_chrom):
    """The on-target flags of BinCounts.LoadBinCounts (CanvasNormalize/BinCounts.cs:102-160): bins in file order, manifest
    regions per chromosome as sorted 1-based (start, end) pairs.  A bin is on target when the first region of its chromosome
    that does not end before the bin's first base (End >= bin.Start + 1) starts at or before its last base
    (Start <= bin.Stop); the region cursor only moves forward and restarts when the chromosome NAME changes.  The result is
    the `on_target` argument of cg_normalize_reference / cg_normalize_best_lr2 / cg_normalize_ratio."""
    import numpy as np
    out = np.zeros(len(chrom), np.uint8)
    cur, regions, idx = None, None, -1
    for i, (c, a, b) in enumerate(zip(np.asarray(chrom).tolist(), np.asarray(start).tolist(), np.asarray(stop).tolist())):
        name = names[c]
        if name != cur:
            cur = name
            regions = regions_by_chrom.get(name)
            if regions is not None:
                idx = 0
        while regions is not None and idx < len(regions) and regions[idx][1] < a + 1:
            idx += 1
        if regions is not None and idx < len(regions) and regions[idx][0] <= b:
            out[i] = 1
    return out


class PloidyInfo:
    """CanvasCommon.PloidyInfo (PloidyInfo.cs:12-178): reference ploidy intervals per chromosome from a ploidy VCF
    (END in INFO, CN in the single genotype column; "." = 2)."""

    def __init__(self):
        self.by_chr = {}  # chromosome -> [(one-based start, one-based end, ploidy)]

    @staticmethod
    def load_vcf_no_sample_id(path):
        """LoadPloidyFromVcfFileNoSampleId (:112-127) + LoadPloidyFromVcfFile (:129-165)."""
        info = PloidyInfo()
        samples = None
        with _open_text(path) as f:
            for line in f:
                line = line.rstrip("\r\n")
                if line.startswith("##") or not line:
                    continue
                if line.startswith("#"):
                    samples = line.split("\t")[9:]
                    if len(samples) == 0:
                        raise ValueError(f"File '{path}' does not contain any genotype column")
                    if len(samples) > 1:
                        raise ValueError(f"File '{path}' cannot have more than one genotype columns when no sample ID provided'")
                    continue
                if samples is None:
                    raise ValueError(f"File '{path}' does not contain any genotype column")
                t = line.split("\t")
                fields = dict(kv.split("=", 1) for kv in t[7].split(";") if "=" in kv)
                genotype = dict(zip(t[8].split(":"), t[9].split(":")))
                if "CN" not in genotype:
                    raise ValueError(f"File '{path}' must contain one genotype CN column!")
                cn = 2 if genotype["CN"] == "." else int(genotype["CN"])
                info.by_chr.setdefault(t[0], []).append((int(t[1]), int(fields["END"]), cn))
        if samples is None:
            raise ValueError(f"File '{path}' does not contain any genotype column")
        return info

    def _counts(self, chrom, one_based_start, one_based_end):
        """getPloidyCounts (:92-109): bases of the query at each ploidy 0..4."""
        counts = [0, 0, one_based_end - one_based_start + 1, 0, 0]
        for a, b, ploidy in self.by_chr[chrom]:
            if ploidy == 2:
                continue
            lo = max(one_based_start - 1, a - 1)
            if lo > b:
                continue
            n = min(one_based_end, b) - lo
            if n <= 0:
                continue
            counts[2] -= n
            counts[ploidy] += n  # IndexError above 4, as the reference's int[5]
        return counts

    def reference_copy_number(self, chrom, begin, end):
        """GetReferenceCopyNumber (:56-72) of a segment [begin, end) in bed coordinates: the ploidy covering most bases,
        the lowest on ties, 2 when nothing is covered."""
        if chrom not in self.by_chr:
            return 2
        best, cn = 0, 2
        for k, v in enumerate(self._counts(chrom, begin + 1, end)):
            if v > best:
                best, cn = v, k
        return cn

    def is_uniform(self, chrom, one_based_start, one_based_end):
        """IsUniformReferencePloidy (:78-90)."""
        if chrom not in self.by_chr:
            return True
        return sum(1 for v in self._counts(chrom, one_based_start, one_based_end) if v > 0) < 2


def post_process_segments(order, seg_by_chr, start, end, cov, excluded=None, max_inter_bin_dist=1000000, ploidy=None):
    """SegmentationResultsProcessor.PostProcessSegments (SegmentationResultsProcessor.cs:17-129); `ploidy` is a
    PloidyInfo or None.  Returns chr -> list of segments {id, bins: [(start, end, coverage)]}."""
    excluded = excluded or {}
    starts = set()
    for c, segs in seg_by_chr.items():
        for a, _ in segs:
            starts.add((c, a))
    seg_num = -1
    out = {}
    for c in order:
        out[c] = []
        cur = None
        ex = excluded.get(c)
        ex_idx = 0
        prev_end = 0
        for a, b, v in zip(start[c].tolist(), end[c].tolist(), cov[c].tolist()):
            new = (c, a) in starts
            if ex is not None:
                while ex_idx < len(ex) and ex[ex_idx][1] < prev_end:
                    ex_idx += 1
                if ex_idx < len(ex):
                    mid = (ex[ex_idx][0] + ex[ex_idx][1]) // 2
                    if prev_end < mid and b >= mid:
                        new = True
            if prev_end > 0 and max_inter_bin_dist >= 0 and prev_end + max_inter_bin_dist < a and not new:
                new = True
            # a change of reference ploidy between the end of the last bin and the end of this one (:116-125)
            if not new and ploidy is not None and not ploidy.is_uniform(c, prev_end if prev_end > 0 else 1, b):
                new = True
            if new:
                seg_num += 1
                cur = {"id": seg_num, "bins": [(a, b, v)]}
                out[c].append(cur)
            elif cur is None:
                cur = {"id": seg_num, "bins": [(a, b, v)]}
                out[c].append(cur)
            else:
                cur["bins"].append((a, b, v))
            prev_end = b
    return out


def segment_extent(seg):
    """SegmentWithBins.Start / End (Models/SegmentWithBins.cs:38-52): smallest start and largest end over the bins, in
    whatever order they were added."""
    return min(a for a, _, _ in seg["bins"]), max(b for _, b, _ in seg["bins"])


def segment_median_coverage(seg):
    """SegmentWithBins.MedianCoverage (Models/SegmentWithBins.cs:18-21): MathNet Statistics.Median of the bins' coverage —
    the mean of the two middle values for an even count (pinned by SegmentWithBinsTests.AddBinTest)."""
    v = sorted(c for _, _, c in seg["bins"])
    if not v:
        return float("nan")
    m = len(v) // 2
    return v[m] if len(v) % 2 else (v[m - 1] + v[m]) / 2.0


def write_partitioned(path, order, segments):
    """SegmentationInput.WriteCanvasPartitionResults (Segmentation.cs:235-252)."""
    buf = io.StringIO()
    text = {}  # coverage values repeat (a few thousand distinct two-decimal numbers): format each once
    for c in order:
        for seg in segments[c]:
            tail = f"\t{seg['id']}\n"
            for a, b, v in sorted(seg["bins"], key=lambda t: t[0]):
                t = text.get(v)
                if t is None:
                    t = text[v] = dotnet_double(v)
                buf.write(f"{c}\t{a}\t{b}\t{t}{tail}")
    with open(path, "wb") as f:
        f.write(gzip_bytes(buf.getvalue().encode()))


def read_cleaned_columns(path):
    """chr, start, stop, count columns of a .cleaned file as parsed by MergeMultiSampleCleanedBedFile
    (Utilities.cs:876-891: int.Parse, float.Parse)."""
    chrom, start, stop, count = [], [], [], []
    with _open_text(path) as f:
        for line in f:
            p = line.rstrip("\n").rstrip("\r").split("\t")
            if len(p) < 4:
                continue
            chrom.append(p[0])
            start.append(int(p[1]))
            stop.append(int(p[2]))
            count.append(float(p[3]))
    return chrom, np.array(start, np.int32), np.array(stop, np.int32), np.array(count, np.float32)


def _runs(chrom):
    """Chromosome names of a column in order of first appearance."""
    out, seen = [], set()
    for c in chrom:
        if c not in seen:
            seen.add(c)
            out.append(c)
    return out


def merge_chromosome_orders(orders):
    """One chromosome order that every file's own order is a subsequence of (when such an order exists): ids assigned from
    it make every sample's (id, start) key column ascending, which cg_merge_common_bins requires.  Numbering by first
    appearance over the files does not: a chromosome that the first file lost completely (chrY filtered out of the mother)
    but a later file still has mid-genome would get the highest id and make that later file non-monotone.  A name not seen
    before is inserted right after the file's previous chromosome."""
    merged = []
    for order in orders:
        at = 0
        for c in order:
            if c in merged:
                at = merged.index(c) + 1
            else:
                merged.insert(at, c)
                at += 1
    return merged


def normalize_canvas_clean(engine, cleaned_paths):
    """CanvasRunner.NormalizeCanvasClean (CanvasRunner.cs:883-903): rewrite every sample's .cleaned file with the bins
    common to all samples — four columns, the count printed by float.ToString().  The set intersection runs on the
    GPU (cg_merge_common_bins); rows must be in (chromosome, start) order, which CanvasClean's output is."""
    cols = [read_cleaned_columns(p) for p in cleaned_paths]
    names = merge_chromosome_orders([_runs(chrom) for chrom, _, _, _ in cols])
    ids = {c: i for i, c in enumerate(names)}
    if len(names) > 256:
        raise ValueError("more than 256 chromosomes (contigs): this build addresses chromosomes with 8-bit ids (DESIGN.md, Limits)")
    samples = [(np.array([ids[c] for c in chrom], np.uint8), a, b, v) for chrom, a, b, v in cols]
    r = engine.merge_common_bins(samples)
    chrom0 = cols[0][0]
    start0 = cols[0][1]
    kept = r["kept_index"]
    for k, path in enumerate(cleaned_paths):
        txt = textcodec.float_default_text(r["count"][k])
        buf = io.StringIO()
        for i, b, t in zip(kept.tolist(), r["stop"].tolist(), txt):
            buf.write(f"{chrom0[i]}\t{int(start0[i])}\t{b}\t{t}\n")
        with open(path, "wb") as f:
            f.write(gzip_bytes(buf.getvalue().encode()))
    return len(kept)
