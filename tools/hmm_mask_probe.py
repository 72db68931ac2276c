"""Probe: PerSampleHMM of one full-size cleaned sample with chromosome masks (the shares of 1, 2, 3 ranks): kernel time, wall
time and whether the call fell back to the sequential kernel (stats[0])."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from canvas_b200 import native, synth

eng = native.Engine(0)
s = synth.make_sample(config=4, sample=0, scale=1.0, n_events=60)
c = eng.clean(s.chrom, s.is_autosome, s.is_chr_y, s.start, s.stop, s.count, s.gc)
off = synth.chrom_offsets(s.chrom[c["kept_index"]], len(s.names))
lens = np.diff(off)
for ranks in (1, 2, 3):
    own = native.shard_assign(lens, ranks)
    for r in range(ranks):
        mask = (own == r).astype(np.uint8)
        for rep in range(3):
            t0 = time.perf_counter()
            out = eng.partition_hmm_counts(off, c["count"], text_mode=2, per_sample=True, chrom_selected=None if ranks == 1 else mask)
            wall = (time.perf_counter() - t0) * 1e3
        st = eng.last_partition_stats_raw()
        print(f"ranks {ranks} share {r}: chromosomes {int(mask.sum())} bins {int(lens[mask > 0].sum())} kernel {out['kernel_ms']:.3f} ms wall {wall:.3f} ms "
              f"sequential {st[0]} blocks {st[1]} launches {out['launches']} stages {eng.last_stage_ms()}")
