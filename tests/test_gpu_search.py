"""-m gpu: batched A* / Theta* / Lazy Theta* vs the canonical oracle — identical node sequences, bit-equal costs."""
import numpy as np
import pytest

from util import assert_results_equal, make_pair

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world(b200, po, synth):
    gm, om, c = make_pair(b200, po, synth, "C3")
    s, g = synth.random_queries(5, om.get_inflate_occupancy, c["origin"], c["size"], 300)
    return gm, om, c, s, g


def _run_both(b200, po, gm, om, algo, s, g, allocate_num=1000000, lam=5.0, res=0.1, semantics="canonical", **kw):
    pl = b200.Planner(algo, resolution=res, lambda_heuristic=lam, allocate_num=allocate_num, semantics=semantics, **kw)
    pl.setGridMap(gm)
    rg, pg = pl.findPaths(s, g, path_cap=len(s) * 4096)
    op = po.OraclePlanner(algo, om, res, lam, allocate_num, mode=po.FAITHFUL if semantics == "faithful" else po.CANONICAL, **kw)
    ro, pth = op.find_paths(s, g, path_stride=4096, nthreads=8)
    return pl, rg, pg, ro, pth


def _assert_paths_equal(b200, rg, pg, ro, pth, stride=4096):
    for i in range(len(rg)):
        n = int(ro["n_path"][i])
        assert rg["n_path"][i] == n
        if n == 0:
            continue
        a = b200.path_of(rg, pg, i)
        assert a is not None
        b = pth[i * stride: i * stride + n]
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), f"query {i}: node sequence differs"


@pytest.mark.parametrize("algo", ["astar", "thetastar", "lazythetastar"])
def test_batched_queries_match_canonical_oracle(b200, po, world, algo):
    gm, om, c, s, g = world
    n = 300 if algo != "thetastar" else 120
    pl, rg, pg, ro, pth = _run_both(b200, po, gm, om, algo, s[:n], g[:n])
    assert ro["solved"].all()
    assert_results_equal(rg, ro)
    _assert_paths_equal(b200, rg, pg, ro, pth)
    assert pl.last_launches() >= 1


@pytest.mark.parametrize("algo", ["astar", "thetastar"])
def test_faithful_semantics_match_the_reference_restatement(b200, po, world, algo):
    """The default semantics of astar/thetastar: bit-identical to the oracle's FAITHFUL mode, which the CPU
    suite pins against the reference's own compiled code (tests/test_oracle_golden.py, test_oracle_vs_ref.py)."""
    gm, om, c, s, g = world
    n = 300 if algo == "astar" else 120
    pl, rg, pg, ro, pth = _run_both(b200, po, gm, om, algo, s[:n], g[:n], semantics="faithful")
    assert ro["solved"].all()
    assert_results_equal(rg, ro)
    _assert_paths_equal(b200, rg, pg, ro, pth)
    # and "default" means faithful for these two
    pl2 = b200.Planner(algo, resolution=0.1, lambda_heuristic=5.0)
    pl2.setGridMap(gm)
    rd, _ = pl2.findPaths(s[:n], g[:n])
    assert np.array_equal(rd["g_final"], rg["g_final"]) and np.array_equal(rd["explored_nodes"], rg["explored_nodes"])


@pytest.mark.parametrize("algo", ["astar", "thetastar"])
def test_faithful_hard_queries_lambda_one(b200, po, synth, algo):
    # lambda = 1: thousands of expansions, many re-keyed open nodes -> the tree-shape-dependent pops matter most here
    gm, om, c = make_pair(b200, po, synth, "C3")
    s, g = synth.random_queries(9, om.get_inflate_occupancy, c["origin"], c["size"], 6 if algo == "astar" else 3, min_dist=8.0)
    pl, rg, pg, ro, pth = _run_both(b200, po, gm, om, algo, s, g, lam=1.0, allocate_num=150000, semantics="faithful")
    assert_results_equal(rg, ro)
    _assert_paths_equal(b200, rg, pg, ro, pth)


def test_faithful_pool_exhaustion_and_sealed_goal(b200, po, world):
    gm, om, c, s, g = world
    for algo in ["astar", "thetastar"]:
        pl, rg, pg, ro, pth = _run_both(b200, po, gm, om, algo, s[:40], g[:40], allocate_num=400, semantics="faithful")
        assert (ro["status"] == 2).any()
        assert_results_equal(rg, ro)


def test_faithful_is_rejected_for_extension_algorithms(b200):
    with pytest.raises(b200.B200Error):
        b200.Planner("lazythetastar", resolution=0.1, semantics="faithful")


@pytest.mark.parametrize("algo", ["astar", "thetastar"])
def test_single_query_api_and_result_object(b200, po, world, algo):
    gm, om, c, s, g = world
    pl = b200.Planner(algo, resolution=0.1, lambda_heuristic=5.0, semantics="canonical")
    pl.setGridMap(gm)
    d = pl.findPath(s[0], g[0])
    op = po.OraclePlanner(algo, om, 0.1, 5.0, 1000000, mode=po.CANONICAL)
    r, p = op.find_path(s[0], g[0])
    assert d["solved"] and r["solved"]
    assert d["explored_nodes"] == r["explored_nodes"]
    assert d["line_of_sight_checks"] == r["los_checks"]
    assert d["g_final_node"] == r["g_final"]
    assert d["path_length"] == r["path_length"]
    assert np.array_equal(pl.getPath(), p)
    assert np.array_equal(d["goal_coords"], p[-1])      # last node, not the requested goal (AlgorithmBase.cpp:130)
    assert np.array_equal(d["start_coords"], s[0])
    assert not pl.detectCollision(s[0])
    assert pl.detectCollision((1e3, 0, 0))              # outside the map counts as collision (-1 -> true)


def test_unknown_algorithm_falls_back_to_astar(b200, po, world):
    gm, om, c, s, g = world
    pl = b200.Planner("astarm1", resolution=0.1)
    pl.setGridMap(gm)
    assert pl.algorithm == "astar"
    ref = b200.Planner("astar", resolution=0.1)
    ref.setGridMap(gm)
    a, _ = pl.findPaths(s[:8], g[:8])
    b, _ = ref.findPaths(s[:8], g[:8])
    assert np.array_equal(a["g_final"], b["g_final"])


def test_pool_exhaustion_and_unreachable_goal(b200, po, synth, world):
    gm, om, c, s, g = world
    # (1) tiny node budget -> "run out of memory" (AStar.cpp:165-168)
    for algo in ["astar", "thetastar", "lazythetastar"]:
        pl, rg, pg, ro, pth = _run_both(b200, po, gm, om, algo, s[:40], g[:40], allocate_num=400)
        assert (ro["status"] == 2).any()
        assert_results_equal(rg, ro)
    # (2) goal sealed inside a hollow box -> open set empties (AStar.cpp:175-178)
    c1 = dict(origin=(-3.2, -3.2, -0.01), size=(6.4, 6.4, 3.2))
    grid = np.zeros((64, 64, 32), np.uint8)
    grid[20:30, 20:30, 5:15] = 1
    grid[22:28, 22:28, 7:13] = 0
    gm2 = b200.GridMap(c1["size"], 0.1, origin=c1["origin"]); gm2.set_inflate(grid)
    om2 = po.OracleMap(c1["origin"], c1["size"], 0.1); om2.set_inflate(grid)
    inside = np.array([[-3.2 + 2.5, -3.2 + 2.5, -0.01 + 1.0]])
    outside = np.array([[2.0, 2.0, 2.0]])
    for algo in ["astar", "thetastar", "lazythetastar"]:
        pl, rg, pg, ro, pth = _run_both(b200, po, gm2, om2, algo, inside, outside)
        assert ro["status"][0] == 1 and not ro["solved"][0]
        assert_results_equal(rg, ro)


def test_big_queries_take_the_second_workspace_tier(b200, po, synth, monkeypatch):
    # lambda = 1 (plain A*) explores far more than 32768 nodes -> tier overflow -> rerun in the big tier (the path of the
    # 128-thread and faithful kernels, and of the 32-thread ones once the promotion slots are used up: forced here)
    monkeypatch.setenv("B200_PROMO_SLOTS", "0")
    gm, om, c = make_pair(b200, po, synth, "C3")
    s, g = synth.random_queries(8, om.get_inflate_occupancy, c["origin"], c["size"], 6, min_dist=12.0)
    pl, rg, pg, ro, pth = _run_both(b200, po, gm, om, "astar", s, g, lam=1.0, allocate_num=200000)
    assert (ro["explored_nodes"] > 16384).any()
    assert_results_equal(rg, ro)
    _assert_paths_equal(b200, rg, pg, ro, pth)
    assert pl.last_launches() == 2


@pytest.mark.parametrize("algo", ["astar", "lazythetastar", "costastar"])
def test_big_queries_move_into_a_promotion_slot_and_go_on(b200, po, synth, algo, monkeypatch):
    """Canonical 32-thread kernels: a query that outgrows the 32 768-node tier is moved (nodes, open list, rebuilt hash) into
    a full-size workspace inside the kernel and continues — one launch, results identical to the oracle; with fewer
    promotion slots than big queries the rest takes the re-run path; a promoted single query still serves getVisitedNodes."""
    gm, om, c = make_pair(b200, po, synth, "C3")
    if algo == "costastar":
        gm.build_edt(1.0); om.build_edt(1.0)
    s, g = synth.random_queries(8, om.get_inflate_occupancy, c["origin"], c["size"], 6, min_dist=12.0)
    kw = dict(lam=1.0, allocate_num=200000, semantics="canonical")
    if algo == "costastar":
        kw["cost_weight"] = 1.0
    pl, rg, pg, ro, pth = _run_both(b200, po, gm, om, algo, s, g, **kw)
    big = int((ro["explored_nodes"] > 32768).sum())
    assert big >= 2
    assert_results_equal(rg, ro)
    _assert_paths_equal(b200, rg, pg, ro, pth)
    assert pl.last_launches() == 1
    # a second batch on the same handle: the promotion slots were left clean
    rg2, pg2 = pl.findPaths(s, g, path_cap=pg.shape[0])
    assert_results_equal(rg2, ro)
    # single promoted query: visited nodes come from the promotion slot
    i = int(np.argmax(ro["explored_nodes"]))
    r = pl.findPath(s[i], g[i])
    assert r["explored_nodes"] == int(ro["explored_nodes"][i]) and pl.last_launches() == 1
    assert len(pl.getVisitedNodes()) == r["explored_nodes"] - 1
    # one promotion slot for `big` >= 2 big queries: the others are re-run in a second launch, same results
    monkeypatch.setenv("B200_PROMO_SLOTS", "1")
    rg3, pg3 = pl.findPaths(s, g, path_cap=pg.shape[0])
    assert_results_equal(rg3, ro)
    _assert_paths_equal(b200, rg3, pg3, ro, pth)
    assert pl.last_launches() == 2
    monkeypatch.setenv("B200_PROMO_SLOTS", "0")
    rg4, _ = pl.findPaths(s, g)
    assert_results_equal(rg4, ro)
    assert pl.last_launches() == 2


def test_c1_theta_star_single_query(b200, po, synth):
    # BASELINE config 1: (0,0,0) -> (40,40,4) on the 250x250x50 grid at 0.2 m
    gm, om, c = make_pair(b200, po, synth, "C1")
    s, g = np.array([[0.0, 0.0, 0.0]]), np.array([[40.0, 40.0, 4.0]])
    assert om.get_inflate_occupancy(s)[0] == 0 and om.get_inflate_occupancy(g)[0] == 0
    for algo in ["thetastar", "astar", "lazythetastar"]:
        pl, rg, pg, ro, pth = _run_both(b200, po, gm, om, algo, s, g, res=0.2)
        assert ro["solved"][0]
        assert_results_equal(rg, ro)
        _assert_paths_equal(b200, rg, pg, ro, pth)


def test_max_los_extension(b200, po, world):
    gm, om, c, s, g = world
    for algo in ["thetastar", "lazythetastar"]:
        pl, rg, pg, ro, pth = _run_both(b200, po, gm, om, algo, s[:40], g[:40], max_los=2.0)
        assert_results_equal(rg, ro)
        _assert_paths_equal(b200, rg, pg, ro, pth)


@pytest.mark.parametrize("algo,semantics", [("thetastaragr", "faithful"), ("thetastaragr", "canonical"), ("thetastaragrpp", "canonical"),
                                            ("thetastaragrpp", "faithful")])
def test_air_ground_planners_match_oracle(b200, po, world, algo, semantics):
    gm, om, c, s, g = world
    s, g = s[:60].copy(), g[:60].copy()
    s[:6, 2] = 0.0          # on the ground plane: rolling start, and AGRpp's 0/0 = NaN edge
    s[6:9, 2] = -0.004      # below it: clamped to 0 (ThetaStarAGR.cpp:141-142)
    g[9:15, 2] = 0.05
    kw = dict(barrier=2.0, flying_cost=10.0, flying_cost_default=0.3, ground_judge=0.1, epsilon=0.1)
    pl, rg, pg, ro, pth = _run_both(b200, po, gm, om, algo, s, g, allocate_num=200000, semantics=semantics, **kw)
    assert ro["solved"].sum() > 40
    assert_results_equal(rg, ro)
    _assert_paths_equal(b200, rg, pg, ro, pth)


@pytest.mark.parametrize("algo,mode", [("astar", "faithful"), ("thetastar", "faithful"), ("lazythetastar", "canonical")])
def test_visited_nodes_match_oracle(b200, po, synth, algo, mode):
    """getVisitedNodes (AlgorithmBase.cpp:185-189) after a single findPath: same nodes in the same creation order."""
    gm, om, c = make_pair(b200, po, synth, "C3")
    s, g = synth.random_queries(9, om.get_inflate_occupancy, c["origin"], c["size"], 3, min_dist=6.0)
    pl = b200.Planner(algo, resolution=c["resolution"], lambda_heuristic=5.0, allocate_num=200000, semantics=mode)
    pl.setGridMap(gm)
    op = po.OraclePlanner(algo, om, c["resolution"], 5.0, 200000, mode=po.FAITHFUL if mode == "faithful" else po.CANONICAL)
    for i in range(len(s)):
        r = pl.findPath(s[i], g[i])
        o, _ = op.find_path(s[i], g[i])
        vg, vo = pl.getVisitedNodes(), op.visited()
        assert r["explored_nodes"] == int(o["explored_nodes"]) and len(vg) == r["explored_nodes"] - 1
        assert np.array_equal(vg, vo)
    pl.findPaths(s, g)
    with pytest.raises(b200.B200Error):
        pl.getVisitedNodes()  # only defined after a single-query call
