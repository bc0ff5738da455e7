"""CPU: hand-derived known-answer facts about the reference's behaviour (SURVEY Appendices A-C) and the
internal consistency of the oracle's own extensions (EDT, cost field, lazy / cost planners)."""
import numpy as np


def small_map(po, dims=(40, 40, 20), res=0.1, origin=(-2.0, -2.0, -0.01)):
    size = tuple(d * res for d in dims)
    return po.OracleMap(origin, size, res)


def test_geometry_and_address_layout(po):
    m = po.OracleMap((-25.6, -25.6, -0.01), (51.2, 51.2, 12.8), 0.1)
    assert m.dims == (512, 512, 128)                      # ceil(size/res) exact for BASELINE sizes (A.1)
    m2 = po.OracleMap((-5, -5, -0.01), (50, 50, 10), 0.2)
    assert m2.dims == (250, 250, 50)


def test_pos_to_index_is_rounding_dependent(po):
    # Appendix C.1: accumulating x += r from a voxel-boundary start lands many sites in voxel k-1
    m = po.OracleMap((-20.0, -20.0, -0.01), (40.0, 40.0, 5.0), 0.1)
    x, low = 0.0, 0
    for k in range(1, 200):
        x += 0.1
        idx = m.pos_to_index([[x, 0.0, 1.0]])[0, 0]
        low += int(idx == 200 + k - 1)
    assert 100 < low < 199


def test_is_in_map_margin(po):
    m = small_map(po)
    o = -2.0
    pts = np.array([[o + 1e-4, 0, 1], [np.nextafter(o + 1e-4, -np.inf), 0, 1], [0, 0, -0.01 + 0.5e-4], [0, 0, 1.99 - 1e-4 - 1e-9]])
    assert list(m.is_in_map(pts)) == [True, False, False, True]
    assert list(m.get_inflate_occupancy(pts)) == [0, -1, -1, 0]


def test_single_point_inflates_3x3x3_in_position_space(po, synth):
    m = small_map(po)
    p = np.array([[0.234, -0.511, 0.777]], np.float32)
    m.update_cloud(synth.pack_cloud(p), (0, 0, 1))
    g = m.inflate()
    assert g.sum() == 27                                    # s = ceil(0.099/0.1) = 1 in x,y; z fixed +-1 (grid_map.cpp:735-736)
    idx = m.pos_to_index(p.astype(np.float64))[0]
    assert g[idx[0] - 1: idx[0] + 2, idx[1] - 1: idx[1] + 2, idx[2] - 1: idx[2] + 2].sum() == 27
    lo, hi = m.local_bounds()
    assert lo[0] <= idx[0] - 1 and hi[0] >= idx[0] + 1


def test_z_inflation_is_always_one_voxel(po, synth):
    m = po.OracleMap((-2.0, -2.0, -0.01), (4.0, 4.0, 2.0), 0.1, obstacles_inflation=0.25)
    m.update_cloud(synth.pack_cloud(np.array([[0.234, -0.511, 0.777]], np.float32)), (0, 0, 1))
    assert m.inflate().sum() == 7 * 7 * 3                   # s = 3 in x,y, still +-1 in z


def test_update_range_filter_is_strict_and_per_axis(po, synth):
    m = po.OracleMap((-2.0, -2.0, -0.01), (4.0, 4.0, 2.0), 0.1, local_update_range=(0.5, 0.5, 0.5))
    pts = np.array([[0.5, 0.0, 1.0], [0.49, 0.0, 1.0], [0.0, 0.0, 1.5]], np.float32)   # |dx| < 0.5 strict
    m.update_cloud(synth.pack_cloud(pts), (0.0, 0.0, 1.0))
    g = m.inflate()
    assert g.sum() == 27


def test_dk_sequence_and_los_semantics(po):
    dk = po.los_dk(1200)
    assert dk[0] == 0.0 and dk[1] == 0.05 and dk[2] == 0.1
    assert (dk != np.arange(1200) * 0.05).sum() > 1000      # A.4: d_k by repeated addition is not k*0.05
    m = small_map(po)
    grid = np.zeros(m.dims, np.uint8)
    grid[20, :, :] = 1                                      # wall at x index 20
    m.set_inflate(grid)
    a, b = np.array([[-1.0, 0.0, 1.0]]), np.array([[1.0, 0.0, 1.0]])
    vis, ns = m.los_batch(a, b, return_samples=True)
    assert vis[0] == 0 and 18 <= ns[0] <= 22                # blocked ~1 m in, at the first 0.05 m sample inside the wall
    vis, ns = m.los_batch(a, a + [[0.8, 0.0, 0.0]], return_samples=True)
    assert vis[0] == 1 and ns[0] == 16                      # d_16 = 0.8000000000000002 > L: the end point is NOT sampled (A.4)
    assert m.los_batch(a, a)[0] == 1                        # zero length: one sample at the start
    assert m.los_batch(a, [[5.0, 0.0, 1.0]])[0] == 0        # leaving the map counts as blocked (-1 is truthy)
    assert m.los_batch([[-5.0, 0.0, 1.0]], a)[0] == 0       # starting outside the map


def test_astar_termination_box_and_path_length_quirk(po):
    m = small_map(po)
    s, g = (-1.0, -1.0, 0.5), (1.03, 0.77, 1.21)
    for mode in (po.FAITHFUL, po.CANONICAL):
        r, p = po.OraclePlanner("astar", m, 0.1, 5.0, 100000, mode=mode).find_path(s, g)
        assert r["solved"] and np.array_equal(p[0], s)      # start is NOT snapped (A.5)
        assert (np.abs(p[-1] - np.array(g)) <= 0.1).all()   # termination box, last node != goal (B6)
        assert np.array_equal(r["goal_coords"], p[-1])
        true_len = np.linalg.norm(np.diff(p, axis=0), axis=1).sum()
        assert abs(r["path_length"] - true_len * 0.1) < 1e-4   # calculatePathLength multiplies by res again (A.7)
        assert r["explored_nodes"] > r["iter_num"]          # nodes created, not expanded


def test_theta_star_shortcuts_in_free_space(po):
    m = small_map(po)
    s, g = (-1.5, -1.5, 0.3), (1.5, 1.2, 1.5)
    ra, pa = po.OraclePlanner("astar", m, 0.1, 5.0, 100000).find_path(s, g)
    rt, pt = po.OraclePlanner("thetastar", m, 0.1, 5.0, 100000).find_path(s, g)
    rl, pl = po.OraclePlanner("lazythetastar", m, 0.1, 5.0, 100000).find_path(s, g)
    assert rt["n_path"] <= 3 and rl["n_path"] <= 4 and ra["n_path"] > 20
    assert rt["g_final"] <= ra["g_final"] + 1e-9 and rl["g_final"] <= ra["g_final"] + 1e-9
    assert rt["los_checks"] > 10 * rl["los_checks"] > 0     # one LoS per neighbour vs one per expansion


def test_edt_lower_envelope_equals_brute_force(po):
    rng = np.random.default_rng(0)
    for dims, dens in [((13, 17, 9), 0.02), ((20, 8, 31), 0.2), ((9, 9, 9), 0.0), ((6, 30, 5), 0.5)]:
        m = po.OracleMap((0, 0, 0), tuple(d * 0.5 for d in dims), 0.5)
        assert m.dims == dims
        m.set_inflate((rng.random(dims) < dens).astype(np.uint8))
        d2, _ = m.build_edt()
        assert np.array_equal(d2, m.edt_brute())


def test_cost_field_definition(po):
    m = po.OracleMap((0, 0, 0), (2.0, 2.0, 2.0), 0.1)
    grid = np.zeros(m.dims, np.uint8)
    grid[10, 10, 10] = 1
    m.set_inflate(grid)
    d2, cq, cf = m.build_edt(d_safe=0.5, want_cost=True)
    assert d2[10, 10, 10] == 0 and cq[10, 10, 10] == 65535 and cf[10, 10, 10] == 1.0
    assert d2[13, 14, 10] == 25 and cq[13, 14, 10] == 0     # 0.5 m away: cost reaches 0 at d_safe
    assert 0 < cq[12, 10, 10] < 65535 and abs(cf[12, 10, 10] - (1 - 0.2 / 0.5)) < 1e-6
    assert (np.diff(cq[10:16, 10, 10].astype(int)) <= 0).all()


def test_cost_planners_prefer_clearance(po, synth):
    m = po.OracleMap((-3.2, -3.2, -0.01), (6.4, 6.4, 3.2), 0.1)
    grid = np.zeros(m.dims, np.uint8)
    grid[30:34, 0:40, :] = 1                                # wall with a gap, path must hug its end
    m.set_inflate(grid)
    m.build_edt(d_safe=1.0)
    s, g = (-2.0, -2.0, 1.0), (2.5, -2.0, 1.0)
    for base in ("astar", "lazythetastar"):
        r0, p0 = po.OraclePlanner(base, m, 0.1, 5.0, 200000).find_path(s, g)
        r1, p1 = po.OraclePlanner("cost" + base, m, 0.1, 5.0, 200000, cost_weight=5.0).find_path(s, g)
        assert r0["solved"] and r1["solved"]
        assert r1["g_final"] > r0["g_final"]                # cost term adds to g


def test_dda_line_of_sight_facts(po, synth):
    """The voxel-traversal LoS: never more permissive than 'every voxel on the exact path is free',
    blocked by a one-voxel wall a coarse sampled march could straddle, and symmetric in nothing it should not be."""
    om = po.OracleMap((-3.2, -3.2, -0.01), (6.4, 6.4, 3.2), 0.1, obstacles_inflation=0.0, local_update_range=(1e3, 1e3, 1e3))
    grid = np.zeros(om.dims, dtype=np.uint8)
    grid[40, :, :] = 1   # a one-voxel-thick wall at x index 40
    grid[40, 30, 10] = 0  # with a single-voxel hole
    om.set_inflate(grid)
    c = lambda i, j, k: ((i + 0.5) * 0.1 - 3.2, (j + 0.5) * 0.1 - 3.2, (k + 0.5) * 0.1 - 0.01)
    vis, cells = om.los_batch_dda([c(20, 30, 10), c(20, 31, 10), c(20, 30, 10), c(20, 30, 10), (100.0, 0.0, 1.0)],
                                  [c(60, 30, 10), c(60, 31, 10), c(30, 50, 20), c(20, 30, 10), c(20, 30, 10)], return_cells=True)
    assert vis.tolist() == [1, 0, 1, 1, 0]
    assert cells[0] == 41 and cells[3] == 1 and cells[4] == 0  # straight along x: 41 voxels; same voxel: 1; outside: none
    # a voxel on the path blocks, whichever end it is at
    grid[20, 30, 10] = 1
    om.set_inflate(grid)
    assert om.los_batch_dda([c(20, 30, 10)], [c(30, 30, 10)]).tolist() == [0]
    assert om.los_batch_dda([c(30, 30, 10)], [c(20, 30, 10)]).tolist() == [0]
