"""CPU research (NOT product code, not used by tests): a FOUR-level hierarchy at 2M poses (coarsest level of 756 rows solved exactly,
i.e. amg_dense_max >= 756) against the five levels the library builds today, for different numbers of inner K-cycle steps.
python tools/research/deep_levels_2m.py [poses]
Result at 2M poses (library today: five levels, three steps everywhere: 62 iterations in this prototype, 66 measured, ~417 coarse
kernels per iteration): four levels with steps (3, 2): 71, (3, 3): 58, (3, 4): 49, (4, 3): 47 -- (3, 3) needs ~129 coarse kernels."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools" / "research"))
from smoothed_level0 import *
import smoothed_level0 as sl
n = int(sys.argv[1])
g, H0, b, pos0 = build(n)
lv = sl.hierarchy(g, H0, pos0, 0, dense_max=1024)
print("levels", [L.H.shape[0] // 3 for L in lv], flush=True)
for ms in ((3, 2), (3, 3), (3, 4), (4, 3)):
    t = time.time()
    its = fcg(H0, b, lambda r: hs.cyc(lv, 0, r, ms), rtol=1e-9)[1]
    print(f"{n} poses K-cycle{ms}: PCG its {its} ({time.time()-t:.0f} s)", flush=True)
