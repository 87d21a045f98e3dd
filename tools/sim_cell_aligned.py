"""What would a CELL-ALIGNED list order buy?  (CPU model, no GPU: same set-up as tools/sim_gather_lines.py)
Today every lane walks its own list front to back, so at iteration r the 32 lanes of a warp gather 32 records that lie in
~23 different 128-byte lines.  Alternative: the warp walks the neighbour CELLS in lock step -- all lanes gather from the
same cell at the same time; the segment of a cell is padded to the longest lane.  Fewer distinct lines per gather (the
records of one cell are contiguous), but more iterations (padding).  Printed per warp: iterations, gather line touches,
lane slots used -- for both orders."""
import numpy as np
from scipy.spatial import cKDTree
a0 = 3.615; nc = 20; rl = 5.9
base = np.array([[0, 0, 0], [.5, .5, 0], [.5, 0, .5], [0, .5, .5]])
g = np.stack(np.meshgrid(*[np.arange(nc)] * 3, indexing='ij'), -1).reshape(-1, 1, 3)
x0 = ((g + base[None]).reshape(-1, 3) * a0) + 0.25 * a0
box = nc * a0; n = len(x0)
gd = int(box / rl); cs = box / gd
rng = np.random.default_rng(1)
for sig in (0.08, 0.12):
    x = (x0 + rng.normal(0, sig, x0.shape)) % box
    c = np.floor(x / cs).astype(int) % gd
    cid = (c[:, 0] * gd + c[:, 1]) * gd + c[:, 2]
    order = np.lexsort((np.arange(n), cid))
    xs = x[order]; cids = cid[order]
    t = cKDTree(xs, boxsize=box)
    nb = t.query_ball_point(xs, rl - 1e-12)
    lists = []
    for i, l in enumerate(nb):
        l = np.array(sorted(v for v in l if v != i))
        d = xs[l] - xs[i]; d -= box * np.round(d / box); r = np.sqrt((d * d).sum(1))
        lists.append(l[r <= 5.5 + 0.25])            # a mid-cycle walk: inner pairs + part of the skin
    it0 = ln0 = sl0 = it1 = ln1 = 0
    W = min(n // 32, 250)
    for w in range(W):
        ls = [lists[i] for i in range(32 * w, 32 * w + 32)]
        m = max(len(l) for l in ls)
        for r in range(m):
            act = np.array([l[r] for l in ls if len(l) > r])
            it0 += 1; ln0 += len(np.unique(act // 4)); sl0 += len(act)
        cells = np.unique(np.concatenate([cids[l] for l in ls]))
        for cc in cells:
            seg = [l[cids[l] == cc] for l in ls]
            m = max(len(s_) for s_ in seg)
            for r in range(m):
                act = np.array([s_[r] for s_ in seg if len(s_) > r])
                it1 += 1; ln1 += len(np.unique(act // 4))
    print(f"sigma {sig:.2f}: own order   {it0 / W:6.1f} iterations/warp, {ln0 / W:7.1f} line touches/warp ({ln0 / it0:4.1f} per gather), {sl0 / W:7.1f} lane slots")
    print(f"            cell-aligned {it1 / W:6.1f} iterations/warp, {ln1 / W:7.1f} line touches/warp ({ln1 / it1:4.1f} per gather)")
