import numpy as np, sys
from scipy.spatial import cKDTree
a0 = 3.615; nc = 16; rl = 5.9
base = np.array([[0,0,0],[.5,.5,0],[.5,0,.5],[0,.5,.5]])
g = np.stack(np.meshgrid(*[np.arange(nc)]*3, indexing='ij'), -1).reshape(-1,1,3)
x0 = ((g + base[None]).reshape(-1,3) * a0) + 0.25*a0
box = nc*a0; n = len(x0)
gd = int(box/rl); cs = box/gd
def morton(q, bits):
    key = np.zeros(len(q), np.int64)
    for b in range(bits-1, -1, -1):
        key = (key << 3) | (((q[:,0]>>b)&1)<<2) | (((q[:,1]>>b)&1)<<1) | ((q[:,2]>>b)&1)
    return key
def run(sig, bits, L):
    rng = np.random.default_rng(1)
    x = (x0 + rng.normal(0, sig, x0.shape)) % box
    c = np.floor(x / cs).astype(int) % gd
    cid = (c[:,0]*gd + c[:,1])*gd + c[:,2]
    num = rng.permutation(n)
    if bits:
        q = np.floor((x/cs - np.floor(x/cs)) * (1<<bits)).astype(int)
        order = np.lexsort((num, morton(q, bits), cid))
    else:
        order = np.lexsort((num, cid))
    xs = x[order]
    t = cKDTree(xs, boxsize=box)
    nb = t.query_ball_point(xs, rl - 1e-12)
    lists = []
    for i, l in enumerate(nb):
        l = np.array(sorted(v for v in l if v != i))
        d = xs[l] - xs[i]; d -= box*np.round(d/box); r = np.sqrt((d*d).sum(1))
        cls = np.clip(np.ceil((r - 5.5)/0.1), 0, 4).astype(int)
        k = np.lexsort((l, cls))
        lists.append(l[k])
    tot_it = tot_lines = tot_pairs = 0
    apw = 32 // L
    for w in range(0, min(n//apw, 400)):
        ls = [lists[i] for i in range(apw*w, apw*w+apw)]
        rows = max((len(l) + L - 1)//L for l in ls)
        for r in range(rows):
            act = np.concatenate([l[r*L:(r+1)*L] for l in ls])
            if len(act) == 0: continue
            tot_it += 1; tot_lines += len(np.unique(act // 4)); tot_pairs += len(act)
    return tot_lines/tot_it, tot_lines/tot_pairs
for sig in (0.0, 0.16):
    for bits in (0, 2):
        print(sig, bits, [tuple(round(v,2) for v in run(sig, bits, L)) for L in (1, 2, 4, 8)])
