"""Prototype (numpy) of the phase-parallel emulation of PropagationDistanceField::propagatePositive: every bucket is processed
in rounds of data-parallel phases whose result equals the serial queue's.  Checked against the oracle's C restatement."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle.df import PropagationDistanceField

DIRS = [(dx, dy, dz) for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1)]  # direction number order
def dirnum(d): return (d[0] + 1) * 9 + (d[1] + 1) * 3 + d[2] + 1
NB = [[[], ] * 27, [[], ] * 27]
for n in range(2):
    NB[n] = []
    for s in DIRS:
        l = []
        for t in DIRS:
            if t == (0, 0, 0): continue
            if n >= 1:
                if abs(t[0]) + abs(t[1]) + abs(t[2]) != 1: continue
                if s[0] * t[0] < 0 or s[1] * t[1] < 0 or s[2] * t[2] < 0: continue
            l.append(t)
        NB[n].append(l)
# slot tables: SLOT[n][dir][dirnum(t)] = index of t in NB[n][dir] or -1
SLOT = np.full((2, 27, 27), -1, np.int64)
for n in range(2):
    for s in range(27):
        for k, t in enumerate(NB[n][s]):
            SLOT[n, s, dirnum(t)] = k
NBARR = np.zeros((2, 27, 26, 3), np.int64)
NBCNT = np.zeros((2, 27), np.int64)
for n in range(2):
    for s in range(27):
        NBCNT[n, s] = len(NB[n][s])
        for k, t in enumerate(NB[n][s]):
            NBARR[n, s, k] = t
DIRARR = np.array(DIRS, np.int64)

class Emul:
    def __init__(self, n, max_d2):
        self.n = n; self.N = n[0] * n[1] * n[2]; self.max = max_d2
        self.dist = np.full(self.N, max_d2, np.int64)
        self.cp = np.full(self.N, -1, np.int64)
        self.dir = np.zeros(self.N, np.int64)
        self.buckets = [[] for _ in range(max_d2 + 1)]
        self.stats = dict(rounds=0, hazards=0, backward=0, accepted=0, entries=0)

    def coords(self, v):
        nz, ny = self.n[2], self.n[1]
        return np.stack([v // (ny * nz), (v // nz) % ny, v % nz], -1)

    def add_voxels(self, vox):  # vox: linear indices in push order (duplicates allowed)
        vox = np.asarray(vox, np.int64)
        self.dist[vox] = 0; self.cp[vox] = vox; self.dir[vox] = 13
        self.buckets[0].append(vox)
        self.propagate()

    def propagate(self):
        n = np.array(self.n); N = self.N
        for i in range(self.max + 1):
            if not self.buckets[i]: continue
            L = np.concatenate(self.buckets[i]); self.buckets[i] = []
            if len(L) == 0: continue
            self.stats["entries"] += len(L)
            d = min(i, 1); S = 26 if d == 0 else 6
            a = 0; end = len(L)
            while a < end:
                self.stats["rounds"] += 1
                R = L[a:end]; P = np.arange(a, end)
                first = np.full(N, -1, np.int64); last = np.full(N, -1, np.int64)
                # first / last position of every voxel in the round's range
                first[R[::-1]] = P[::-1]; last[R] = P
                cv = self.coords(self.cp[R]); dv = self.dir[R]; xv = self.coords(R)
                # all (p, slot) offers
                pp, ss = np.meshgrid(np.arange(len(R)), np.arange(S), indexing="ij")
                pp = pp.ravel(); ss = ss.ravel()
                ok = ss < NBCNT[d, dv[pp]]
                pp, ss = pp[ok], ss[ok]
                diff = NBARR[d, dv[pp], ss]
                tc = xv[pp] + diff
                ok = ((tc >= 0) & (tc < n)).all(1)
                pp, ss, diff, tc = pp[ok], ss[ok], diff[ok], tc[ok]
                newd = ((cv[pp] - tc) ** 2).sum(1)
                ok = newd <= self.max
                pp, ss, diff, tc, newd = pp[ok], ss[ok], diff[ok], tc[ok], newd[ok]
                t = (tc[:, 0] * n[1] + tc[:, 1]) * n[2] + tc[:, 2]
                cand = newd < self.dist[t]
                pp, ss, diff, tc, newd, t = pp[cand], ss[cand], diff[cand], tc[cand], newd[cand], t[cand]
                key = (pp + a) * S + ss
                acc = np.ones(len(t), bool)
                # earlier offers to the same target from the other entries of the round
                for du in (DIRARR if d == 0 else DIRARR[[4, 10, 12, 14, 16, 22]]):
                    if (du == 0).all(): continue
                    uc = tc - du
                    okb = ((uc >= 0) & (uc < n)).all(1)
                    u = (uc[:, 0] * n[1] + uc[:, 1]) * n[2] + uc[:, 2]
                    u = np.where(okb, u, 0)
                    fp = np.where(okb, first[u], -1)
                    has = fp >= 0
                    slot = SLOT[d, self.dir[u], dirnum(du)]
                    has &= slot >= 0
                    cu = self.coords(self.cp[u])
                    nd_u = ((cu - tc) ** 2).sum(1)
                    key_u = fp * S + slot
                    acc &= ~(has & (nd_u <= self.max) & (key_u < key) & (nd_u <= newd))
                pp, ss, diff, newd, t, key = pp[acc], ss[acc], diff[acc], newd[acc], t[acc], key[acc]
                # hazards: an accepted offer changes a voxel that still has an entry later in this round
                hz = last[t] > pp + a
                p_h = end
                if hz.any():
                    self.stats["hazards"] += int(hz.sum())
                    q = np.where(first[t] > pp + a, first[t], pp + a + 1)[hz]
                    p_h = int(q.min())
                keep = pp + a < p_h
                pp, ss, diff, newd, t, key = pp[keep], ss[keep], diff[keep], newd[keep], t[keep], key[keep]
                self.stats["accepted"] += len(t)
                self.stats["backward"] += int((newd <= i).sum())
                # apply: the final state of a target is its accepted offer of smallest distance
                np.minimum.at(self.dist, t, newd)
                fin = self.dist[t] == newd
                self.cp[t[fin]] = self.cp[R[pp[fin]]]
                self.dir[t[fin]] = (diff[fin, 0] + 1) * 9 + (diff[fin, 1] + 1) * 3 + diff[fin, 2] + 1
                # pushes in key order, stable by destination bucket
                fw = newd != i  # == i: pushed behind the captured end and cleared; < i: stays queued for the NEXT propagation
                o = np.argsort(key[fw], kind="stable")
                tk, nk = t[fw][o], newd[fw][o]
                o2 = np.argsort(nk, kind="stable")
                tk, nk = tk[o2], nk[o2]
                if len(nk):
                    b, st = np.unique(nk, return_index=True)
                    st = list(st) + [len(nk)]
                    for j, bb in enumerate(b):
                        self.buckets[bb].append(tk[st[j]:st[j + 1]])
                a = p_h


def run(seed, n_cells, res, maxd, n_pts, calls=1, clustered=False):
    rng = np.random.default_rng(seed)
    size = [c * res for c in n_cells]
    f = PropagationDistanceField(size[0], size[1], size[2], res, 0, 0, 0, maxd)
    e = Emul(f.num_cells, f.max_distance_sq)
    for c in range(calls):
        if clustered:
            ctr = rng.uniform(0, 1, (max(1, n_pts // 50), 3)) * size
            pts = (ctr[:, None, :] + rng.normal(0, 2 * res, (len(ctr), 50, 3))).reshape(-1, 3)
        else:
            pts = rng.uniform(-0.1, 1.1, (n_pts, 3)) * size
        # addPointsToField: valid points whose voxel is not an obstacle yet, input order, duplicates kept
        cells = np.floor(pts / res + 0.0).astype(np.int64)  # origin 0; the oracle's own worldToGrid decides below
        vox = []
        for p in pts:
            c3, ok = f.world_to_grid(*p)
            if ok:
                v = (c3[0] * f.num_cells[1] + c3[1]) * f.num_cells[2] + c3[2]
                if e.dist[v] > 0:
                    vox.append(v)
        f.add_points(pts)
        e.add_voxels(vox)
        ref = f.d2().ravel()
        bad = int((ref != e.dist).sum())
        print("seed %d call %d: %d voxels, %d obstacles, mismatches %d, stats %s" % (seed, c, e.N, int((ref == 0).sum()), bad, e.stats))
        if bad: return False
    return True

if __name__ == "__main__":
    ok = True
    ok &= run(1, (20, 20, 20), 0.05, 0.3, 30)
    ok &= run(2, (30, 25, 20), 0.02, 0.2, 100, calls=3)
    ok &= run(3, (50, 50, 50), 0.02, 0.2, 2000)
    ok &= run(4, (40, 40, 40), 0.02, 0.3, 20, calls=4)
    ok &= run(5, (40, 40, 40), 0.02, 0.3, 300, calls=3, clustered=True)
    print("ALL OK" if ok else "FAILED")
