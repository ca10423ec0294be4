"""Extra randomized parity sweep on the GPU (beyond the seeds of tests/): python tools/parity_sweep.py LO HI"""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_cases as pc
from transcriptclean_b200.engine import Engine
lo, hi = int(sys.argv[1]), int(sys.argv[2])
eng = Engine(0)
bad = 0
t0 = time.time()
for seed in range(lo, hi):
    kw = {}
    if seed % 5 == 0: kw = dict(err=0.10, shift_frac=0.5)
    if seed % 7 == 0: kw = dict(exon_len=(12, 60), err=0.04)
    if seed % 11 == 0: kw = dict(tagless_frac=1.0)
    if seed % 13 == 0: eng.set_pipeline(29)
    try:
        pc.check_generated(eng, seed, n_reads=120, unpinned=(seed % 3 == 0), **kw)
    except Exception as e:
        bad += 1
        print("SEED", seed, kw, "FAILED:", str(e)[:1500], flush=True)
    finally:
        eng.set_pipeline(131072)
print("swept %d seeds in %.0f s, %d failures" % (hi - lo, time.time() - t0, bad))
