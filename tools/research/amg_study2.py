import sys, time
sys.path.insert(0, "tools/research")
from amg_proto import *
n = int(sys.argv[1])
g, H, b, pos = build(n)
xd = spla.splu(H.tocsc()).solve(b)
for name, kw, cyc in [
    ("two-grid chain4", dict(max_levels=2), dict()),
    ("two-grid graph16", dict(max_levels=2, first="graph"), dict()),
    ("three-level chain4", dict(max_levels=3), dict()),
    ("two-grid SA graph16", dict(max_levels=2, first="graph", smooth_P=True), dict()),
    ("two-grid chain4 V(2,2)", dict(max_levels=2), dict(nu=2)),
]:
    t = time.time(); lv = setup(H, pos, **kw); ts = time.time() - t
    sizes = [L.H.shape[0] // 3 for L in lv]
    t = time.time(); x, it = pcg(H, b, lambda r: vcycle(lv, 0, r, **cyc)); tp = time.time() - t
    print(f"{name:30s} its {it:5d} err {np.linalg.norm(x-xd)/np.linalg.norm(xd):.1e} levels {sizes} setup {ts:.1f}s pcg {tp:.1f}s", flush=True)
