"""Offline model: distinct 128-byte lines per warp gather for candidate memory orders / list layouts."""
import numpy as np, sys
from scipy.spatial import cKDTree
rng = np.random.default_rng(1)
a0 = 3.615; nc = 24; rl = 5.9
base = np.array([[0,0,0],[.5,.5,0],[.5,0,.5],[0,.5,.5]])
g = np.stack(np.meshgrid(*[np.arange(nc)]*3, indexing='ij'), -1).reshape(-1,1,3)
x = ((g + base[None]).reshape(-1,3) * a0)
box = nc*a0
sig = float(sys.argv[1]) if len(sys.argv) > 1 else 0.08
x = (x + rng.normal(0, sig, x.shape)) % box
n = len(x)
gd = int(box/rl); cs = box/gd
print("atoms", n, "cells", gd, "cell size", cs)

def neighbors(order):
    """order: permutation new->old. returns positions in new order and list of neighbor arrays (new indices)"""
    xs = x[order]
    t = cKDTree(xs, boxsize=box)
    nb = t.query_ball_point(xs, rl - 1e-12)
    return xs, [np.array(sorted(v for v in l if v != i)) for i, l in enumerate(nb)]

def stats(name, lists, warp_rows):
    """warp_rows(w) -> 2D int array [iters][32] of j or -1"""
    tot_it = 0; tot_lines = 0; tot_act = 0; nw = 0
    for w in range(0, min(n//32, 600)):
        rows = warp_rows(w)
        for r in rows:
            act = r[r >= 0]
            if len(act) == 0: continue
            tot_it += 1; tot_act += len(act); tot_lines += len(np.unique(act >> 2))
        nw += 1
    print(f"{name:34s} iters/warp {tot_it/nw:6.1f}  active/iter {tot_act/tot_it:5.1f}  lines/iter {tot_lines/tot_it:5.1f}  lines/warp {tot_lines/nw:7.1f}")

def plain_rows(lists):
    def f(w):
        ls = [lists[i] for i in range(32*w, 32*w+32)]
        m = max(len(l) for l in ls)
        R = -np.ones((m, 32), int)
        for k, l in enumerate(ls): R[:len(l), k] = l
        return R
    return f

def padded_rows(lists, winid):
    """pad per window: winid[j] -> window id of atom j (monotone in j)"""
    def f(w):
        ls = [lists[i] for i in range(32*w, 32*w+32)]
        wins = sorted(set(np.concatenate([winid[l] for l in ls])))
        out = []
        for wd in wins:
            sub = [l[winid[l] == wd] for l in ls]
            m = max(len(s) for s in sub)
            R = -np.ones((m, 32), int)
            for k, s in enumerate(sub): R[:len(s), k] = s
            out.append(R)
        return np.concatenate(out)
    return f

# S0: cell-major z fastest, in-cell by atom number
c = np.floor(x / cs).astype(int) % gd
cid = (c[:,0]*gd + c[:,1])*gd + c[:,2]
o0 = np.lexsort((np.arange(n), cid))
xs, L0 = neighbors(o0)
print("mean nbrs", np.mean([len(l) for l in L0]))
stats("S0 cell-major, by number", L0, plain_rows(L0))
cid_s = cid[o0]
stats("S3 cell-major, pad per cell", L0, padded_rows(L0, cid_s))
stats("S4 cell-major, pad per z-run(col)", L0, padded_rows(L0, cid_s // gd))
# S0b: in-cell sorted by z
o = np.lexsort((x[:,2], cid)); xs, L = neighbors(o)
stats("S0b cell-major, in-cell by z", L, plain_rows(L))
for nsub in (2, 3, 4):
    # pencils: fine grid in x,y (nsub per cell), whole column along z, sorted by z
    f = np.floor(x[:, :2] / (cs/nsub)).astype(int) % (gd*nsub)
    pid = f[:,0]*gd*nsub + f[:,1]
    o = np.lexsort((x[:,2], pid)); xs, L = neighbors(o)
    stats(f"S1 pencils {nsub}x{nsub}, z-sorted", L, plain_rows(L))
    stats(f"S2 pencils {nsub}x{nsub}, pad per pencil", L, padded_rows(L, pid[o]))

print("---- direction/shell sorted lists ----")
def resort(xs, lists, qr, qd):
    out = []
    for i, l in enumerate(lists):
        d = xs[l] - xs[i]; d -= box*np.round(d/box)
        r = np.sqrt((d*d).sum(1))
        kb = np.floor(r/qr).astype(int) if qr > 0 else np.zeros(len(l), int)
        q = np.floor(d/qd + 0.5).astype(int)
        key = np.lexsort((q[:,2], q[:,1], q[:,0], kb))
        out.append(l[key])
    return out
def run(name, o, qr=0.3, qd=1.0):
    xs, L = neighbors(o)
    stats(name + " j-sorted", L, plain_rows(L))
    L2 = resort(xs, L, qr, qd)
    stats(name + f" shell{qr}/dir{qd}", L2, plain_rows(L2))
    L3 = resort(xs, L, 0, qd)
    stats(name + f" dir{qd} only", L3, plain_rows(L3))
run("S0", o0)
for nsub in (3, 4, 5):
    f = np.floor(x[:, :2] / (cs/nsub)).astype(int) % (gd*nsub)
    pid = f[:,0]*gd*nsub + f[:,1]
    run(f"pencil{nsub}", np.lexsort((x[:,2], pid)))
# cell-major, in-cell fine order (sub-pencils inside the cell, z-sorted inside)
for nsub in (3, 4):
    f = np.floor((x[:, :2] % cs) / (cs/nsub)).astype(int)
    key = cid*nsub*nsub + f[:,0]*nsub + f[:,1]
    run(f"cell+subpencil{nsub}", np.lexsort((x[:,2], key)))
