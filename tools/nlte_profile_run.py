"""Two MALI iterations of the CaII problem on NCOL columns (for ncu captures)."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
from pyrh_b200 import nlte
from pyrh_b200.api import Context
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
fixture = sys.argv[2] if len(sys.argv) > 2 else "nlte_caii"
g = dict(np.load(f'/root/repo/tests/golden/{fixture}.npz'))
ctx = Context(0)
prob = nlte.NlteProblem.from_golden(g, ncol=ncol)
out = nlte.iterate(ctx, prob, nmax=2, limit=0.0)
print(out["niter"][:3])
