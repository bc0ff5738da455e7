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


def _sparse_grid(rng, dims, n_occ, empty_planes=()):
    grid = np.zeros(dims, dtype=np.uint8)
    idx = tuple(rng.integers(0, d, n_occ) for d in dims)
    grid[idx] = 1
    for ax, i in empty_planes:
        sl = [slice(None)] * 3
        sl[ax] = i
        grid[tuple(sl)] = 0
    return grid


# every tile shape / entry width the fused kernels instantiate: (dims, what it exercises)
FUSED_SHAPES = [
    ((96, 130, 50), "small family: 8 bands x 32 lanes, 16-bit entries, packed 16-bit intermediate, odd band lengths"),
    ((40, 64, 128), "small family, 4 full z words, bulk-copied plane"),
    ((1024, 60, 40), "large family: pass 2 lines of 1024 (16 x 16 tile), 32-bit (s, g) entries"),
    ((24, 1500, 40), "large family: pass 1 lines of 1500 (16 x 16 tile), own-word staging"),
    ((20, 3000, 33), "large family: pass 1 lines of 3000 (32 x 8 tile)"),
    ((3000, 20, 33), "large family: pass 2 lines of 3000 (32 x 8 tile)"),
    ((3, 2, 5), "degenerate: fewer sites than bands"),
]


@pytest.mark.parametrize("dims,what", FUSED_SHAPES, ids=[str(d) for d, _ in FUSED_SHAPES])
def test_fused_edt_equals_hbm_stack_pipeline_and_oracle(b200, po, dims, what):
    """The two-kernel shared-memory EDT and the seven-kernel HBM-stack EDT are independent implementations (convex-hull
    stack without separators vs Meijster's integer separators); both must equal the CPU lower envelope bit for bit."""
    rng = np.random.default_rng(sum(dims))
    res = 0.5
    size = tuple(d * res for d in dims)
    for n_occ, empties in [(max(2, int(np.prod(dims)) // 4000), ((0, dims[0] // 2), (1, 0))), (1, ()), (0, ())]:
        grid = _sparse_grid(rng, dims, n_occ, empties)
        gm = b200.GridMap(size, res, origin=(0, 0, 0))
        gm.set_inflate(grid)
        gm.build_edt(d_safe=4.0)
        assert gm.edt_pipeline_used() == "fused", what
        d2, cq = gm.dist2().copy(), gm.costq().copy()
        gm.set_edt_pipeline(1)
        gm.build_edt(d_safe=4.0)
        assert gm.edt_pipeline_used() == "hbm-stack"
        assert np.array_equal(d2, gm.dist2()) and np.array_equal(cq, gm.costq()), what
        om = po.OracleMap((0, 0, 0), size, res)
        om.set_inflate(grid)
        d2o, cqo = om.build_edt(d_safe=4.0, nthreads=8)
        assert np.array_equal(d2, d2o) and np.array_equal(cq, cqo), what
        if n_occ == 0:
            assert (d2 == 1 << 29).all() and (cq == 0).all()


def test_fused_edt_slab_with_extent_hint(b200):
    """z-slab build with the total-height hint (32-bit hull entries in pass 2) == without it == the unsharded transform."""
    from astar_planners_b200 import shard
    rng = np.random.default_rng(11)
    dims, splits, res = (600, 36, 96), (32, 64), 0.5
    grid = _sparse_grid(rng, dims, 300)
    grid[:, :, 30:60] = 0
    full = b200.GridMap(tuple(d * res for d in dims), res, origin=(0, 0, 0))
    full.set_inflate(grid)
    full.build_edt(d_safe=3.0)
    want = full.dist2()
    for hint in (0, dims[2]):
        slabs, z0 = [], 0
        for nz in splits:
            m = b200.GridMap((dims[0] * res, dims[1] * res, nz * res), res, origin=(0, 0, z0 * res))
            m.set_inflate(np.ascontiguousarray(grid[:, :, z0:z0 + nz]))
            m.set_slab_extent(hint)
            slabs.append(m)
            z0 += nz
        ext = [m.column_extents() for m in slabs]
        got = []
        for r, m in enumerate(slabs):
            below, above = shard.slab_gaps([e[0] for e in ext], [e[1] for e in ext], list(splits), r)
            m.build_edt_slab(3.0, below if r > 0 else None, above if r < len(slabs) - 1 else None)
            assert m.edt_pipeline_used() == "fused"
            got.append(m.dist2())
        assert np.array_equal(np.concatenate(got, axis=2), want)


def test_distance_field_is_invalidated_by_grid_writes(b200, po, synth):
    gm, om, c = make_pair(b200, po, synth, "C3")
    gm.build_edt(d_safe=1.0)
    assert gm.has_distance_field(1.0) and not gm.has_distance_field(2.0)
    gm.setOccupied([(0.0, 0.0, 1.0)])
    assert not gm.has_distance_field(1.0)
    with pytest.raises(b200.B200Error):
        gm.dist2()
    pl = b200.Planner("costastar", resolution=0.1, cost_weight=1.0)
    pl.setGridMap(gm)
    with pytest.raises(b200.B200Error):
        pl.findPaths(np.zeros((1, 3)), np.ones((1, 3)))
    gm.build_edt(d_safe=1.0)
    assert gm.getInflateOccupancy([(0.0, 0.0, 1.0)])[0] == 1 and gm.dist2()[tuple(gm.posToIndex([(0.0, 0.0, 1.0)])[0])] == 0
    for mutate in (lambda: gm.resetBuffer(), lambda: gm.cloudCallback(synth.pack_cloud(synth.config_cloud("C3")), (0.0, 0.0, 1.0)),
                   lambda: gm.set_inflate(gm.inflate())):
        gm.build_edt(d_safe=1.0)
        mutate()
        assert not gm.has_distance_field(1.0)


def test_slab_gap_kernel_matches_host_derivation(b200):
    """b200_map_column_extents16 + b200_map_slab_gaps (device path of the C5 exchange) == shard.slab_gaps on the host."""
    import torch
    from astar_planners_b200 import shard
    rng = np.random.default_rng(3)
    dims, splits, res = (50, 37, 160), (32, 64, 32, 32), 0.5
    grid = _sparse_grid(rng, dims, 500)
    grid[:, :, 40:120] *= (rng.random((dims[0], dims[1], 1)) < 0.1)   # most columns empty across two whole slabs
    slabs, z0 = [], 0
    for nz in splits:
        m = b200.GridMap((dims[0] * res, dims[1] * res, nz * res), res, origin=(0, 0, z0 * res))
        m.set_inflate(np.ascontiguousarray(grid[:, :, z0:z0 + nz]))
        slabs.append(m)
        z0 += nz
    ext = [m.column_extents() for m in slabs]
    packed = [m.column_extents16() for m in slabs]
    for e, p in zip(ext, packed):
        assert np.array_equal(p[..., 0], e[0]) and np.array_equal(p[..., 1], e[1])
    gathered = torch.from_numpy(np.stack(packed)).cuda()
    for r, m in enumerate(slabs):
        below = torch.empty(dims[:2], dtype=torch.int32, device="cuda")
        above = torch.empty(dims[:2], dtype=torch.int32, device="cuda")
        m.slab_gaps(gathered, splits, r, below, above)
        m.sync()
        wb, wa = shard.slab_gaps([e[0] for e in ext], [e[1] for e in ext], list(splits), r)
        assert np.array_equal(below.cpu().numpy(), wb) and np.array_equal(above.cpu().numpy(), wa)
