"""Randomised parity fuzz of the +-0 streaming kernel (zero_span_kernel) against the oracle:
random frame sizes incl. partial right/bottom blocks, batches of 1..3 pairs.
usage: python tools/r0_fuzz.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import motionestimation_b200 as me  # noqa: E402
from oracle_binding import Oracle  # noqa: E402

orc=Oracle(); rng=np.random.Generator(np.random.PCG64(5)); bad=0
for i in range(120):
    B=int(rng.choice([8,16])); W=int(rng.integers(B,300)); H=int(rng.integers(B,200))
    if rng.random()<0.5: W=(W//B)*B
    if rng.random()<0.5: H=(H//B)*B
    cur,ref=me.random_pair(W,H,int(rng.integers(1<<30)))
    npairs=int(rng.integers(1,4))
    c=np.stack([cur]*npairs); r=np.stack([ref]*npairs)
    with me.Estimator(W,H,B,0,max_pairs=npairs) as est:
        out=est.search_u8(c,r); k=est.kernel_in_use
    exp=orc.search(cur,ref,B,0)
    for p in range(npairs):
        ok=np.array_equal(out['mvx'][p],exp['mvx']) and np.array_equal(out['mvy'][p],exp['mvy']) and np.array_equal(out['ssd'][p],exp['ssd']) and np.array_equal(out['score'][p].view(np.uint32),exp['score'].view(np.uint32))
        if not ok: bad+=1; print('MISMATCH',W,H,B,p,k)
print('r0 fuzz: 120 cases', bad, 'mismatches')
