"""Synthetic WGS-shaped coverage arrays.

Bins follow hg19 chromosome lengths at 1 kb; counts are Poisson(100 * gc_bias * cn / 2) — the bin
geometry CanvasBin targets (100 counts/bin, reference GermlineWgsRunner.cs:20) — with GC ~ N(41, 6)
autocorrelated over ~25 bins, ~40 copy-number events per genome, 0.1 % single-bin spikes and 2.5 %
oversized bins so every CanvasClean filter fires.  Deterministic: numpy PCG64 seeded by
20240601 + 1000 * config + sample.
"""
import numpy as np

# hg19 chromosome lengths (bp)
HG19 = [
    ("chr1", 249250621), ("chr2", 243199373), ("chr3", 198022430), ("chr4", 191154276),
    ("chr5", 180915260), ("chr6", 171115067), ("chr7", 159138663), ("chr8", 146364022),
    ("chr9", 141213431), ("chr10", 135534747), ("chr11", 135006516), ("chr12", 133851895),
    ("chr13", 115169878), ("chr14", 107349540), ("chr15", 102531392), ("chr16", 90354753),
    ("chr17", 81195210), ("chr18", 78077248), ("chr19", 59128983), ("chr20", 63025520),
    ("chr21", 48129895), ("chr22", 51304566), ("chrX", 155270560), ("chrY", 59373566),
]


def is_autosome(name):
    """Stand-in for Isas GenomeMetadata.SequenceMetadata.IsAutosome [EXT, unpinned]: a name that,
    after an optional 'chr' prefix, parses as an integer."""
    s = name[3:] if name.lower().startswith("chr") else name
    return s.isdigit()


def is_chr_y(name):
    """LoessGCNormalizer.cs:52-53."""
    return name.lower() in ("chry", "y")


class Sample:
    """SoA view of one .binned file (reference SampleGenomicBin, GenomicBin.cs:45-117)."""

    def __init__(self, names, chrom, start, stop, count, gc):
        self.names = list(names)
        self.chrom = np.ascontiguousarray(chrom, np.uint8)
        self.start = np.ascontiguousarray(start, np.int32)
        self.stop = np.ascontiguousarray(stop, np.int32)
        self.count = np.ascontiguousarray(count, np.float32)
        self.gc = np.ascontiguousarray(gc, np.uint8)
        self.is_autosome = np.array([is_autosome(n) for n in self.names], np.uint8)
        self.is_chr_y = np.array([is_chr_y(n) for n in self.names], np.uint8)

    def __len__(self):
        return len(self.count)


def make_sample(config=2, sample=0, chromosomes=None, bin_size=1000, scale=1.0, n_events=40,
                mean_count=100.0, tumour=False):
    """One synthetic .binned array.  `scale` shrinks every chromosome (tests); `chromosomes`
    restricts to a subset of names (config 1 = ['chr20'])."""
    rng = np.random.default_rng(20240601 + 1000 * config + sample)
    # config 4 is a pedigree: its samples are binned on ONE layout (CanvasRunner.cs:846-870 hands the same bin definitions to
    # every sample), so coordinates and GC content come from a generator the sample index does not touch
    lrng = np.random.default_rng(20240601 + 1000 * config + 999) if config == 4 else rng
    chroms = [(n, l) for n, l in HG19 if chromosomes is None or n in chromosomes]
    names = [n for n, _ in chroms]
    cols = {k: [] for k in ("chrom", "start", "stop", "count", "gc")}
    for ci, (name, length) in enumerate(chroms):
        nb = max(30, int(length / bin_size * scale))
        start = np.arange(nb, dtype=np.int64) * bin_size
        stop = start + bin_size
        # 2.5 % oversized bins around the chromosome middle (centromere-like)
        n_big = max(1, int(0.025 * nb))
        mid = nb // 2
        big = np.arange(mid - n_big // 2, mid - n_big // 2 + n_big)
        big = big[(big >= 0) & (big < nb)]
        extra = np.exp(lrng.uniform(np.log(5e3), np.log(3e6), size=len(big))).astype(np.int64)
        shift = np.zeros(nb, np.int64)
        shift[big] = extra
        cum = np.cumsum(shift)
        start = start + cum - shift
        stop = stop + cum
        # GC: autocorrelated N(41, 6)
        white = lrng.normal(0, 1, nb + 24)
        kernel = np.ones(25) / np.sqrt(25)
        smooth = np.convolve(white, kernel, mode="valid")[:nb]
        gcv = 41 + 6 * (0.8 * smooth + 0.6 * lrng.normal(0, 1, nb))
        gc = np.clip(np.rint(gcv), 0, 100).astype(np.int64)
        bias = np.clip(1 - 8e-4 * (gc - 45.0) ** 2, 0.3, 1.1)
        cn = np.full(nb, 2.0 if is_autosome(name) else 1.0)
        n_ev = rng.poisson(max(0.5, n_events * nb / 3.1e6))
        for _ in range(n_ev):
            ln = int(np.exp(rng.uniform(np.log(10), np.log(5000))))
            ln = min(ln, max(1, nb // 4))
            s = int(rng.integers(0, max(1, nb - ln)))
            val = float(rng.choice([0, 1, 3, 4]))
            if tumour:
                val = 2 + 0.7 * (val - 2) * float(rng.choice([1.0, 0.5]))
            cn[s:s + ln] = val
        lam = mean_count * bias * cn / 2
        if tumour:
            r = 50.0
            count = rng.negative_binomial(r, r / (r + np.maximum(lam, 1e-9)))
        else:
            count = rng.poisson(lam)
        spikes = rng.random(nb) < 1e-3
        count = np.where(spikes, count * 3, count)
        cols["chrom"].append(np.full(nb, ci))
        cols["start"].append(start)
        cols["stop"].append(stop)
        cols["count"].append(count)
        cols["gc"].append(gc)
    cat = {k: np.concatenate(v) for k, v in cols.items()}
    return Sample(names, cat["chrom"], cat["start"], cat["stop"], cat["count"], cat["gc"])


def chrom_offsets(chrom, n_chrom):
    """Offsets [n_chrom + 1] of each chromosome in an array sorted by chromosome id."""
    counts = np.bincount(np.asarray(chrom, np.int64), minlength=n_chrom)
    off = np.zeros(n_chrom + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    return off
