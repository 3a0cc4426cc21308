import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qvnt_b200 import workloads, QReg
from tests.test_multi_gpu import run_sharded
n = int(sys.argv[1]); world = int(sys.argv[2]); depth = int(sys.argv[3])
circ = workloads.random_layered(n, depth)
one = QReg.new(n); one.apply(circ); a = one.amplitudes(); one.close()
for rep in range(int(sys.argv[4]) if len(sys.argv) > 4 else 1):
    got, _, _ = run_sharded(n, world, 0, circ, fuse=1)
    nrm = float(np.vdot(got, got).real)
    bad = np.nonzero(np.abs(a - got) > 1e-9)[0]
    print(f"harness n={n} world={world} depth={depth}: norm-1 = {nrm-1:.3e}  max diff {np.abs(a-got).max():.3e} bad {bad.size}", flush=True)
    if bad.size:
        print("  first/last bad index", hex(int(bad[0])), hex(int(bad[-1])))
        orr = int(np.bitwise_or.reduce(bad)); andd = int(np.bitwise_and.reduce(bad))
        print("  OR of bad indices ", bin(orr), "\n  AND of bad indices", bin(andd))
        runs = np.split(bad, np.nonzero(np.diff(bad) != 1)[0] + 1)
        print("  contiguous runs:", len(runs), "lengths", sorted({len(r) for r in runs})[:10], "first run starts", [hex(int(r[0])) for r in runs[:8]])
