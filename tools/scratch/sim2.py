import numpy as np, sys
from scipy.spatial import cKDTree
rng = np.random.default_rng(1)
a0 = 3.615; nc = 24; rl = 5.9
base = np.array([[0,0,0],[.5,.5,0],[.5,0,.5],[0,.5,.5]])
g = np.stack(np.meshgrid(*[np.arange(nc)]*3, indexing='ij'), -1).reshape(-1,1,3)
x = ((g + base[None]).reshape(-1,3) * a0) + 0.25*a0
box = nc*a0
sig = 0.08
x = (x + rng.normal(0, sig, x.shape)) % box
n = len(x)
gd = int(box/rl); cs = box/gd
def neighbors(order):
    xs = x[order]
    t = cKDTree(xs, boxsize=box)
    nb = t.query_ball_point(xs, rl - 1e-12)
    return xs, [np.array(sorted(v for v in l if v != i)) for i, l in enumerate(nb)]
def stats(name, lists, recsz=4):
    tot_it = tot_lines = 0; nw = 0
    for w in range(0, min(n//32, 500)):
        ls = [lists[i] for i in range(32*w, 32*w+32)]
        m = max(len(l) for l in ls)
        for r in range(m):
            act = np.array([l[r] for l in ls if len(l) > r])
            tot_it += 1; tot_lines += len(np.unique(act // recsz))
        nw += 1
    print(f"{name:40s} lines/iter {tot_lines/tot_it:5.1f}")
c = np.floor(x / cs).astype(int) % gd
cid = (c[:,0]*gd + c[:,1])*gd + c[:,2]
xs, L = neighbors(np.lexsort((np.arange(n), cid))); stats("cell-major (current)", L)
for nsx, nsy in ((4,4),(6,6),(8,8),(4,2),(3,3)):
    f = np.stack([np.floor(x[:,0]/(cs/nsx)), np.floor(x[:,1]/(cs/nsy))],1).astype(int)
    pid = f[:,0]*gd*nsy + f[:,1]
    xs, L = neighbors(np.lexsort((x[:,2], pid))); stats(f"pencil {nsx}x{nsy} z-sorted", L)
    # sub-cell version: order by (pid, cell z, atom number)
    xs, L = neighbors(np.lexsort((np.arange(n), c[:,2], pid))); stats(f"subcell {nsx}x{nsy}x1 by number", L)
print("--- hybrid: pencil storage order, list entries in reference-cell scan order ---")
nsx = nsy = 4
f = np.stack([np.floor(x[:,0]/(cs/nsx)), np.floor(x[:,1]/(cs/nsy))],1).astype(int)
pid = f[:,0]*gd*nsy + f[:,1]
order = np.lexsort((np.arange(n), c[:,2], pid))      # storage order (subcell 4x4x1)
xs, L = neighbors(order)
cs_ = np.floor(xs / cs).astype(int) % gd             # reference cell coords in storage order
def scan_key(i, l):
    # order entries as the build scans them: neighbour cell (l,m,n) relative to atom's cell, then storage index
    d = (cs_[l] - cs_[i] + gd//2) % gd - gd//2
    return np.lexsort((l, d[:,2], d[:,1], d[:,0]))
L2 = [l[scan_key(i, l)] for i, l in enumerate(L)]
stats("pencil storage, j-sorted lists", L)
stats("pencil storage, cell-scan-order lists", L2)
# distance-class grouping: class0 (r<=5.5) first then the rest, each j-sorted
def cls_sorted(i, l):
    d = xs[l] - xs[i]; d -= box*np.round(d/box); r = np.sqrt((d*d).sum(1))
    k = np.lexsort((l, r > 5.5))
    return l[k]
L3 = [cls_sorted(i, l) for i, l in enumerate(L)]
stats("pencil storage, class0 first then skin", L3)
