import numpy as np, sys
from scipy.spatial import cKDTree
a0 = 3.615; nc = 20; rl = 5.9
base = np.array([[0,0,0],[.5,.5,0],[.5,0,.5],[0,.5,.5]])
g = np.stack(np.meshgrid(*[np.arange(nc)]*3, indexing='ij'), -1).reshape(-1,1,3)
x0 = ((g + base[None]).reshape(-1,3) * a0) + 0.25*a0
box = nc*a0
n = len(x0)
gd = int(box/rl); cs = box/gd
def run(sig):
    rng = np.random.default_rng(1)
    x = (x0 + rng.normal(0, sig, x0.shape)) % box
    c = np.floor(x / cs).astype(int) % gd
    cid = (c[:,0]*gd + c[:,1])*gd + c[:,2]
    def neighbors(order):
        xs = x[order]
        t = cKDTree(xs, boxsize=box)
        nb = t.query_ball_point(xs, rl - 1e-12)
        out = []
        for i, l in enumerate(nb):
            l = np.array(sorted(v for v in l if v != i))
            d = xs[l] - xs[i]; d -= box*np.round(d/box); r = np.sqrt((d*d).sum(1))
            k = np.lexsort((l, r > 5.5))     # class 0 first, then skin, each j-sorted
            out.append(l[k])
        return out
    def stats(lists):
        tot_it = tot_lines = 0
        for w in range(0, min(n//32, 300)):
            ls = [lists[i] for i in range(32*w, 32*w+32)]
            m = max(len(l) for l in ls)
            for r in range(m):
                act = np.array([l[r] for l in ls if len(l) > r])
                tot_it += 1; tot_lines += len(np.unique(act // 4))
        return tot_lines/tot_it
    a = stats(neighbors(np.lexsort((np.arange(n), cid))))
    f = np.stack([np.floor(x[:,0]/(cs/4)), np.floor(x[:,1]/(cs/4))],1).astype(int)
    pid = f[:,0]*gd*4 + f[:,1]
    b = stats(neighbors(np.lexsort((np.arange(n), c[:,2], pid))))
    print(f"sigma {sig:.2f}: cell-major {a:5.1f}  pencil {b:5.1f}")
for sig in (0.0, 0.04, 0.08, 0.12, 0.16, 0.2):
    run(sig)
