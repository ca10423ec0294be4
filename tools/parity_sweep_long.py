"""Randomized parity sweep with long, error-rich, multi-junction reads (many segments / long lists)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_cases as pc
from transcriptclean_b200.engine import Engine
lo, hi = int(sys.argv[1]), int(sys.argv[2])
eng = Engine(0)
bad = 0
t0 = time.time()
for seed in range(lo, hi):
    kw = dict(exon_len=(150, 700), err=0.08 + 0.01 * (seed % 6), shift_frac=0.6, chrom_len=300000, genes_per_chrom=8)
    try:
        pc.check_generated(eng, seed, n_reads=150, unpinned=(seed % 2 == 0), **kw)
    except Exception as e:
        bad += 1
        print("SEED", seed, kw, "FAILED:", str(e)[:1500], flush=True)
print("swept %d seeds in %.0f s, %d failures" % (hi - lo, time.time() - t0, bad))
