"""CPU (-m "not gpu"): SECOND SOURCES for the builder-defined parts (the reference has no Lazy Theta*,
no cost-aware edge and no EDT, so `oracle.cpp` would otherwise be the only statement of them).

* Lazy Theta*: a pure-Python search written from Nash, Koenig & Tovey, "Lazy Theta*: Any-Angle Path Planning and Path
  Length Analysis in 3D" (AAAI 2010), Algorithm 3 — own open list (heapq with stale-entry skipping), own closed set, own
  line-of-sight loop — on this code base's conventions (FP64 accumulated positions, 26-neighbour stencil from the literal
  loop, fastLOS sampling, tie-break (f, creation index)).  The paper leaves the argmin tie in SetVertex open; with the
  oracle's rule (`via` first, strict <) every field must be bit-identical, with the opposite rule (`via` last, so it loses
  every tie) the path cost must still agree.
* costastar: the same skeleton with the edge  len + w * len * (q_a + q_b) / 131070.
* squared EDT: scipy.ndimage.distance_transform_edt (an independent exact Euclidean transform) squared and rounded.
"""
import heapq
import math

import numpy as np
import pytest

RES, LAM, TIE, LOS_STEP = 0.1, 5.0, 1.0 + 1.0 / 10000, 0.1 / 2.0


class World:
    def __init__(self, origin, size, grid, costq=None):
        self.o, self.grid, self.costq = origin, grid, costq
        self.n = grid.shape
        self.inv = 1 / RES
        self.lo = [origin[i] + 1e-4 for i in range(3)]
        self.hi = [origin[i] + size[i] - 1e-4 for i in range(3)]

    def in_map(self, p):
        return all(self.lo[i] <= p[i] <= self.hi[i] for i in range(3))

    def idx(self, p):
        return tuple(int(math.floor((p[i] - self.o[i]) * self.inv)) for i in range(3))

    def occ(self, p):
        return -1 if not self.in_map(p) else int(self.grid[self.idx(p)])

    def q(self, p):
        i = self.idx(p)
        return float(self.costq[tuple(min(max(i[k], 0), self.n[k] - 1) for k in range(3))])


def norm3(x, y, z):
    return math.sqrt((x * x + y * y) + z * z)


def dist(a, b):
    return norm3(b[0] - a[0], b[1] - a[1], b[2] - a[2])


def line_of_sight(w, a, b, counter):
    """LineOfSight.cpp:9-27 restated independently: d = 0, 0.05, 0.05 + 0.05, ... while d <= |b - a|."""
    v = (b[0] - a[0], b[1] - a[1], b[2] - a[2])
    n = norm3(*v)
    u = (v[0] / n, v[1] / n, v[2] / n) if (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2] > 0 else v
    d = 0.0
    while d <= n:
        counter[0] += 1
        if w.occ((a[0] + u[0] * d, a[1] + u[1] * d, a[2] + u[2] * d)):
            return False
        d += LOS_STEP
    return True


def stencil():
    vals, d = [], -RES
    while d <= RES + 1e-3:
        vals.append(d)
        d += RES
    return [(a, b, c) for a in vals for b in vals for c in vals if not norm3(a, b, c) < 1e-3]


def search(w, start, goal, algo, cost_weight=0.0, via_first=True):
    """algo: 'lazythetastar' | 'astar' | 'costastar'.  Returns dict(solved, g, explored, iters, los_checks, path)."""
    h = lambda p: LAM * (TIE * dist(p, goal))
    edge = (lambda a, b, ln: ln + cost_weight * (ln * ((w.q(a) + w.q(b)) / 131070.0))) if algo == "costastar" else (lambda a, b, ln: ln)
    pos, g, f, parent, via, closed, in_open = [start], [0.0], [h(start)], [None], [None], [False], [True]
    via_len = [0.0]  # c(via, s): the lattice step that generated s from via (not |pos difference|: accumulated positions round)
    index = {start: 0}
    heap = [(f[0], 0)]
    nbrs = stencil()
    los_checks, samples, iters = 0, [0], 0
    while heap:
        fk, s = heapq.heappop(heap)
        if closed[s] or not in_open[s] or fk != f[s]:
            continue  # stale entry
        in_open[s] = False
        if algo == "lazythetastar" and parent[s] is not None and parent[s] != via[s]:  # SetVertex (Algorithm 3, lines 33-37)
            los_checks += 1
            if not line_of_sight(w, pos[parent[s]], pos[s], samples):
                best, best_g = None, math.inf
                if via_first:
                    best, best_g = via[s], g[via[s]] + via_len[s]
                for d in nbrs:
                    t = index.get((pos[s][0] + d[0], pos[s][1] + d[1], pos[s][2] + d[2]))
                    if t is not None and closed[t]:
                        c = g[t] + norm3(*d)
                        if c < best_g:
                            best, best_g = t, c
                if not via_first and g[via[s]] + via_len[s] < best_g:
                    # via IS a closed neighbour, but pos + (-d) need not reproduce its accumulated position bit for bit,
                    # so the exact-key lookups can miss it: the alternative rule takes it last (it loses every tie)
                    best, best_g = via[s], g[via[s]] + via_len[s]
                parent[s], g[s] = best, best_g
                f[s] = best_g + h(pos[s])
        if all(abs(pos[s][i] - goal[i]) <= RES for i in range(3)):
            path, c = [], s
            while c is not None:
                path.append(pos[c])
                c = parent[c]
            return dict(solved=True, g=g[s], explored=len(pos), iters=iters, los_checks=los_checks, path=path[::-1], samples=samples[0])
        closed[s] = True
        iters += 1
        for d in nbrs:
            p = (pos[s][0] + d[0], pos[s][1] + d[1], pos[s][2] + d[2])
            if not w.in_map(p):
                continue
            t = index.get(p)
            if t is not None and closed[t]:
                continue
            if w.occ(p) == 1:
                continue
            if t is None:
                t = len(pos)
                index[p] = t
                pos.append(p); g.append(math.inf); f.append(math.inf); parent.append(None); via.append(None); via_len.append(0.0)
                closed.append(False); in_open.append(False)
            if algo == "lazythetastar" and parent[s] is not None:  # ComputeCost, path 2 assumed (lines 25-31)
                par, ln = parent[s], dist(pos[parent[s]], p)
            else:
                par, ln = s, norm3(*d)
            cand = g[par] + edge(pos[par], p, ln)
            if cand < g[t]:
                g[t], parent[t], via[t], via_len[t] = cand, par, s, norm3(*d)
                f[t] = cand + h(p)
                in_open[t] = True
                heapq.heappush(heap, (f[t], t))
    return dict(solved=False, g=None, explored=len(pos), iters=iters, los_checks=los_checks, path=[], samples=samples[0])


@pytest.fixture(scope="module")
def world(po, synth):
    origin, size = (-1.6, -1.6, -0.01), (3.2, 3.2, 3.2)
    rng = np.random.default_rng(12)
    pts = []
    for _ in range(14):  # short walls and posts
        c = origin + rng.random(3) * size
        ext = rng.choice([(0.9, 0.1, 1.2), (0.1, 0.9, 1.2), (0.2, 0.2, 2.0)])
        pts.append(c + (rng.random((400, 3)) - 0.5) * ext)
    cloud = synth.pack_cloud(np.concatenate(pts).astype(np.float32))
    om = po.OracleMap(origin, size, RES, 0.099, (1e3, 1e3, 1e3), -0.01)
    om.update_cloud(cloud, (0.0, 0.0, 1.0))
    _, cq = om.build_edt(0.5, nthreads=2)
    s, g = synth.random_queries(3, om.get_inflate_occupancy, origin, size, 120, min_dist=1.5, margin=0.15)
    return om, World(origin, size, om.inflate(), cq), s, g


def test_lazy_theta_star_second_source(po, world):
    om, w, s, g = world
    ro, paths = po.OraclePlanner("lazythetastar", om, RES, LAM, 100000, mode=po.CANONICAL).find_paths(s, g, path_stride=512, nthreads=4)
    assert ro["solved"].mean() > 0.9 and ro["los_checks"].sum() > 1000
    same_cost_plain, worst = 0, 0.0
    for i in range(len(s)):
        r = search(w, tuple(s[i]), tuple(g[i]), "lazythetastar")
        assert r["solved"] == bool(ro["solved"][i])
        assert r["explored"] == ro["explored_nodes"][i] and r["iters"] == ro["iter_num"][i], i
        assert r["los_checks"] == ro["los_checks"][i] and r["samples"] == ro["los_samples"][i], i
        if r["solved"]:
            assert r["g"] == ro["g_final"][i], i
            assert np.array_equal(np.array(r["path"]), paths[i * 512: i * 512 + ro["n_path"][i]]), i
        p = search(w, tuple(s[i]), tuple(g[i]), "lazythetastar", via_first=False)  # the paper's argmin, first minimum wins
        assert p["solved"] == r["solved"]
        if p["solved"]:
            # every any-angle segment (longer than one lattice step, i.e. a verified path-2 parent) must be visible,
            # whatever the tie rule; single steps are A*-style moves that neither source tests
            assert all(line_of_sight(w, a, b, [0]) for a, b in zip(p["path"][:-1], p["path"][1:]) if dist(a, b) > 1.8 * RES)
            same_cost_plain += abs(p["g"] - r["g"]) <= 1e-9 * max(1.0, r["g"])
            worst = max(worst, abs(p["g"] - r["g"]) / r["g"])
    # the tie rule only matters where two closed neighbours give the same path-1 cost (measured: all 120 costs identical)
    print("plain-argmin rule: identical cost on", same_cost_plain, "of", int(ro["solved"].sum()), "worst relative deviation", worst)
    assert same_cost_plain >= 0.95 * ro["solved"].sum() and worst < 0.02


@pytest.mark.parametrize("algo,cw", [("astar", 0.0), ("costastar", 3.0)])
def test_astar_and_cost_edge_second_source(po, world, algo, cw):
    om, w, s, g = world
    ro, paths = po.OraclePlanner(algo, om, RES, LAM, 100000, mode=po.CANONICAL, cost_weight=cw).find_paths(s[:60], g[:60], path_stride=512, nthreads=4)
    for i in range(60):
        r = search(w, tuple(s[i]), tuple(g[i]), algo, cost_weight=cw)
        assert r["solved"] == bool(ro["solved"][i]) and r["explored"] == ro["explored_nodes"][i] and r["iters"] == ro["iter_num"][i], i
        if r["solved"]:
            assert r["g"] == ro["g_final"][i]
            assert np.array_equal(np.array(r["path"]), paths[i * 512: i * 512 + ro["n_path"][i]])
    if algo == "costastar":  # the cost term must have been exercised
        r0, _ = po.OraclePlanner("astar", om, RES, LAM, 100000, mode=po.CANONICAL).find_paths(s[:60], g[:60], nthreads=4)
        assert (ro["g_final"] > r0["g_final"]).mean() > 0.5


def test_squared_edt_second_source_scipy(po):
    ndi = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(8)
    for dims, dens in [((40, 33, 21), 0.01), ((64, 64, 32), 0.0005), ((17, 90, 9), 0.05)]:
        grid = (rng.random(dims) < dens).astype(np.uint8)
        grid[0, 0, 0] = 1
        om = po.OracleMap((0, 0, 0), tuple(d * 0.5 for d in dims), 0.5)
        om.set_inflate(grid)
        d2, _ = om.build_edt(0.0, nthreads=2)
        want = np.rint(ndi.distance_transform_edt(grid == 0) ** 2).astype(np.int64)
        assert np.array_equal(d2.astype(np.int64), want)
        assert np.array_equal(om.edt_brute(), d2)
