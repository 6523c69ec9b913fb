"""CPU research (NOT product code, not used by tests): which levels make the PCG count grow with the graph?  The hierarchy of the library
(level-0 aggregates from a structure-only handle, root+neighbours below) with an EXACT solve from a given level size down.
python tools/research/deep_exact.py [poses]
Result (K-cycle PCG iterations at 1M / 2M poses): library hierarchy 44 / 62; exact below level 2: 41 / 44; exact level 1 (two-grid): 37 / 37
=> the level-0 coarse space is size-robust, the growth comes from the inexact solves of levels >= 2."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools" / "research"))
from smoothed_level0 import *
import smoothed_level0 as sl
n = int(sys.argv[1])
g, H0, b, pos0 = build(n)
for dm in (640, 4000, 8000, 130000):
    t = time.time()
    lv = sl.hierarchy(g, H0, pos0, 0, dense_max=dm)
    ms = (3, 2) if len(lv) <= 4 else (3, 3, 3, 3, 3)
    its = fcg(H0, b, lambda r: hs.cyc(lv, 0, r, ms), rtol=1e-9)[1]
    print(f"{n} poses, exact solve from {dm} rows down: levels {[L.H.shape[0] // 3 for L in lv]} K-cycle{ms[:len(lv)-1]} PCG its {its} ({time.time()-t:.0f} s)", flush=True)
