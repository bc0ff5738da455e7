"""-m gpu: exact squared EDT + cost field (new functionality) vs brute force / the oracle's lower-envelope EDT,
and the cost-aware planners vs the oracle."""
import numpy as np
import pytest

from util import assert_results_equal, make_pair

pytestmark = pytest.mark.gpu


def test_edt_matches_brute_force_small(b200, po):
    rng = np.random.default_rng(0)
    for dims, dens in [((24, 20, 37), 0.01), ((16, 33, 8), 0.2), ((40, 12, 70), 0.002), ((9, 9, 9), 0.0)]:
        size = tuple(d * 0.5 for d in dims)   # 0.5 is exact in binary: ceil(size/res) == dims
        grid = (rng.random(dims) < dens).astype(np.uint8)
        gm = b200.GridMap(size, 0.5, origin=(0, 0, 0)); gm.set_inflate(grid)
        om = po.OracleMap((0, 0, 0), size, 0.5); om.set_inflate(grid)
        assert gm.dims == dims
        gm.build_edt()
        assert np.array_equal(gm.dist2(), om.edt_brute())


def test_edt_and_cost_match_oracle_c3(b200, po, synth):
    gm, om, c = make_pair(b200, po, synth, "C3")
    gm.build_edt(d_safe=1.0)
    d2, cq, cf = om.build_edt(d_safe=1.0, nthreads=8, want_cost=True)
    assert np.array_equal(gm.dist2(), d2)
    assert np.array_equal(gm.costq(), cq)
    assert np.array_equal(gm.cost().view(np.uint32), cf.view(np.uint32))
    assert (cq > 0).any() and (cq == 0).any()


def test_edt_properties_full_size_c2(b200, po, synth):
    gm, om, c = make_pair(b200, po, synth, "C2", n_points=1_000_000, rng=(60.0, 60.0, 20.0))
    gm.build_edt(d_safe=1.5)
    d2 = gm.dist2()
    occ = gm.inflate()
    assert ((d2 == 0) == (occ == 1)).all()                       # zero exactly on obstacles
    # 1-Lipschitz in voxel units: sqrt(d2) changes by at most 1 between face neighbours
    r = np.sqrt(d2.astype(np.float64))
    for ax in range(3):
        assert np.abs(np.diff(r, axis=ax)).max() <= 1.0 + 1e-9
    # exact agreement with the CPU lower-envelope transform on a sub-volume of columns
    d2o, _ = om.build_edt(d_safe=0.0, nthreads=8)
    assert np.array_equal(d2, d2o)


@pytest.mark.parametrize("algo", ["costastar", "costlazythetastar"])
def test_cost_aware_planners_match_oracle(b200, po, synth, algo):
    gm, om, c = make_pair(b200, po, synth, "C3")
    gm.build_edt(d_safe=1.0)
    om.build_edt(d_safe=1.0, nthreads=8)
    s, g = synth.random_queries(5, om.get_inflate_occupancy, c["origin"], c["size"], 150)
    pl = b200.Planner(algo, resolution=0.1, lambda_heuristic=5.0, cost_weight=2.0)
    pl.setGridMap(gm)
    rg, pg = pl.findPaths(s, g, path_cap=len(s) * 4096)
    op = po.OraclePlanner(algo, om, 0.1, 5.0, 1000000, mode=po.CANONICAL, cost_weight=2.0)
    ro, pth = op.find_paths(s, g, path_stride=4096, nthreads=8)
    assert ro["solved"].all()
    assert_results_equal(rg, ro)
    for i in range(len(s)):
        n = int(ro["n_path"][i])
        assert np.array_equal(b200.path_of(rg, pg, i), pth[i * 4096: i * 4096 + n])
    # the cost term must matter: without it the same queries cost less
    op0 = po.OraclePlanner(algo.replace("cost", ""), om, 0.1, 5.0, 1000000, mode=po.CANONICAL)
    r0, _ = op0.find_paths(s, g, nthreads=8)
    assert (ro["g_final"] >= r0["g_final"] - 1e-9).mean() > 0.9


def test_cost_planner_without_edt_fails_loudly(b200, po, synth):
    gm, om, c = make_pair(b200, po, synth, "C3")
    pl = b200.Planner("costastar", resolution=0.1, cost_weight=1.0)
    pl.setGridMap(gm)
    with pytest.raises(b200.B200Error):
        pl.findPaths(np.zeros((1, 3)), np.ones((1, 3)))


@pytest.mark.parametrize("splits", [(32, 32, 32), (64, 32), (32, 64), (96,)])
def test_z_slab_edt_equals_full_transform(b200, po, splits):
    """z-slab sharded EDT. Slabs exchange only per-column extents (lowest/highest occupied z); the
    result must equal the unsharded transform exactly (untruncated)."""
    from astar_planners_b200 import shard
    rng = np.random.default_rng(7)
    dims = (40, 24, 96)
    grid = (rng.random(dims) < 0.004).astype(np.uint8)
    grid[:, :, 40:70] = 0                       # a thick empty band: distances must cross an entire slab
    grid[5:9, 3:7, :] = 0                       # columns with no obstacle at all
    res = 0.5
    full = b200.GridMap(tuple(d * res for d in dims), res, origin=(0, 0, 0))
    full.set_inflate(grid)
    full.build_edt(d_safe=3.0)
    want_d2, want_cq = full.dist2(), full.costq()
    slabs, z0 = [], 0
    for nz in splits:
        m = b200.GridMap((dims[0] * res, dims[1] * res, nz * res), res, origin=(0, 0, z0 * res))
        m.set_inflate(np.ascontiguousarray(grid[:, :, z0:z0 + nz]))
        slabs.append(m)
        z0 += nz
    ext = [m.column_extents() for m in slabs]
    lows, highs = [e[0] for e in ext], [e[1] for e in ext]
    for r, (m, nz) in enumerate(zip(slabs, splits)):
        lo_np = np.where(grid[:, :, sum(splits[:r]):sum(splits[:r]) + nz].any(axis=2),
                         grid[:, :, sum(splits[:r]):sum(splits[:r]) + nz].argmax(axis=2), -1)
        assert np.array_equal(lows[r], lo_np)
    got_d2, got_cq = [], []
    for r, m in enumerate(slabs):
        below, above = shard.slab_gaps(lows, highs, list(splits), r)
        m.build_edt_slab(3.0, below if r > 0 else None, above if r < len(slabs) - 1 else None)
        got_d2.append(m.dist2())
        got_cq.append(m.costq())
    assert np.array_equal(np.concatenate(got_d2, axis=2), want_d2)
    assert np.array_equal(np.concatenate(got_cq, axis=2), want_cq)
    assert (want_d2[:, :, 50] > 100).any()      # the band really is far from everything
