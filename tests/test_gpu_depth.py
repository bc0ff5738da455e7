"""Depth-image fusion on the GPU through the C-ABI: bit-exact against the reference's own outputs
(tests/golden/ref_depth_golden.npz) and against the oracle on larger frames."""
import numpy as np
import pytest

from depth_util import GOLDEN, SCENARIOS, check_publishers, replay, scenario

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def _gpu_map(b200, golden, sc):
    gm = b200.GridMap(tuple(golden["size"]), float(golden["res"]), origin=tuple(golden["origin"]), obstacles_inflation=sc["inflation"],
                      local_update_range=sc["range"], ground_height=float(golden["origin"][2]))
    gm.set_depth_params(**sc["params"])
    return gm


@pytest.mark.parametrize("name", SCENARIOS)
def test_depth_fusion_is_bit_identical_to_the_reference(b200, golden, name):
    sc = scenario(golden, name)
    gm = _gpu_map(b200, golden, sc)
    replay(sc, gm.depthPoseCallback, gm.occupancy, gm.inflate, gm.local_bounds, gm.dims, gm.set_depth_frame_count)
    st = gm.depth_stats()
    assert st["projected"] == sc["proj"][-1] and st["frames"] == sc["frames_after"]
    check_publishers(sc, gm.publishMap(), gm.publishUnknown())
    before = gm.occupancy()
    assert not gm.depthPoseCallback(sc["img"][0], (0.0, 0.0, float(golden["size"][2]) + 1.0), sc["rot"][0])  # camera outside: dropped
    assert np.array_equal(before, gm.occupancy())


def test_large_frames_match_the_oracle(b200, po, synth):
    """640x480 frames with the reference's default intrinsics on the C3 grid: grids, bounds and the work counters."""
    c = synth.CONFIGS["C3"]
    rng_ = (5.5, 5.5, 4.5)
    gm = b200.GridMap(c["size"], c["resolution"], origin=c["origin"], local_update_range=rng_, ground_height=c["origin"][2])
    om = po.OracleMap(c["origin"], c["size"], c["resolution"], local_update_range=rng_, ground_height=c["origin"][2])
    gm.set_depth_params()
    om.set_depth_params()
    dp = b200.DEPTH_DEFAULTS
    boxes = synth.pillar_boxes(c["seed"], c["origin"], c["size"], c["n_pillars"])
    r = np.random.default_rng(17)
    base, yaw = np.array([-3.0, 2.0, 1.2]), 0.7
    rounds = []
    for f in range(10):
        pos = base + np.array([0.12 * f, -0.05 * f, 0.0]) + r.uniform(-0.03, 0.03, 3)
        R, _ = synth.look_at(yaw + 0.03 * f, 0.15)
        img = synth.render_depth(boxes, c["origin"][2], pos, R, 480, 640, dp["fx"], dp["fy"], dp["cx"], dp["cy"], dropout=0.02, seed=f)
        assert gm.depthPoseCallback(img, pos, R) == om.update_depth(img, pos, R)
        so, sg = om.depth_stats(), gm.depth_stats()
        assert (sg["projected"], sg["rays_cast"], sg["ray_steps"], sg["updates"]) == tuple(int(x) for x in so[:4]), f"frame {f}"
        rounds.append(sg["order_rounds"])
        if f in (1, 5, 9):
            assert np.array_equal(gm.occupancy(), om.occupancy()), f"frame {f}: log-odds"
            assert np.array_equal(gm.inflate(), om.inflate()), f"frame {f}: inflated grid"
            assert gm.local_bounds() == om.local_bounds()
    occ = om.occupancy()
    assert int((occ > np.log(0.8 / 0.2)).sum()) > 500  # surfaces were seen often enough to become occupied
    assert 2 <= max(rounds) <= 200, rounds  # 35-40 measured on these frames
    # getOccupancy + planning on the fused map
    idx = np.argwhere(occ > np.log(0.8 / 0.2))[:2000]
    pos = (idx + 0.5) * c["resolution"] + np.array(c["origin"])
    q = np.concatenate([pos, pos + [0.0, 0.0, 50.0], r.uniform(-12, 12, (2000, 3))])
    assert np.array_equal(gm.getOccupancy(q), om.get_occupancy(q))
    inside = q[gm.isInMap(q)]
    ii = gm.posToIndex(inside)
    assert np.array_equal(gm.log_odds(inside), occ[ii[:, 0], ii[:, 1], ii[:, 2]])
    assert gm.isUnknown(inside).sum() > 0 and (~gm.isUnknown(inside)).sum() > 0
    # publishMap / publishUnknown blobs (grid_map.cpp:806-849, :902-939)
    pm, pu = gm.publishMap(), gm.publishUnknown()
    assert len(pm) > 500 and len(pu) > 1000
    assert np.array_equal(pm, om.export_occupancy_cloud(0, dp["local_map_margin"] // 2))
    assert np.array_equal(pu, om.export_occupancy_cloud(1, 0))
    pl = b200.Planner("thetastar", resolution=c["resolution"], allocate_num=200000)
    pl.setGridMap(gm)
    op = po.OraclePlanner("thetastar", om, c["resolution"], 5.0, 200000, po.FAITHFUL)
    s, g = (float(base[0]) - 0.5, float(base[1]), 1.0), (float(base[0]) + 4.5, float(base[1]) - 1.0, 1.2)
    a, (b, _) = pl.findPath(s, g), op.find_path(s, g)
    assert a["solved"] == bool(b["solved"]) and a["explored_nodes"] == b["explored_nodes"] and a["g_final_node"] == b["g_final"]


def test_device_image_and_cloud_interplay(b200, po, synth, golden):
    """A depth frame given as a CUDA tensor; a point-cloud frame afterwards writes into the same inflated grid."""
    import torch
    sc = scenario(golden, "nofilter_infl2_edge")
    gm = _gpu_map(b200, golden, sc)
    om = po.OracleMap(tuple(golden["origin"]), tuple(golden["size"]), float(golden["res"]), sc["inflation"], sc["range"],
                      float(golden["origin"][2]))
    om.set_depth_params(**sc["params"])
    for f in range(3):
        img = torch.from_numpy(sc["img"][f].view(np.int16).copy()).cuda()  # same 16 bits; only the pointer is used
        assert gm.depthPoseCallback(img, sc["pos"][f], sc["rot"][f])
        om.update_depth(sc["img"][f], sc["pos"][f], sc["rot"][f])
    cloud = synth.uniform_cloud(3, tuple(golden["origin"]), tuple(golden["size"]), 3000)
    blob = synth.pack_cloud(cloud, 16)
    gm.cloudCallback(blob, sc["pos"][2])
    om.update_cloud(blob, sc["pos"][2])
    assert np.array_equal(gm.inflate(), om.inflate())
    assert np.array_equal(gm.occupancy(), om.occupancy())


def test_depth_errors(b200, golden):
    sc = scenario(golden, "filter")
    gm = b200.GridMap(tuple(golden["size"]), float(golden["res"]), origin=tuple(golden["origin"]))
    with pytest.raises(b200.B200Error):
        gm.depthPoseCallback(sc["img"][0], sc["pos"][0], sc["rot"][0])  # parameters not set
    with pytest.raises(b200.B200Error):
        gm.set_depth_params(p_hit=1.5)
    gm.set_depth_params(depth_filter_margin=0, skip_pixel=2)
    odd = np.ascontiguousarray(sc["img"][0][:119])  # last sampled row = last image row
    gm.depthPoseCallback(odd, sc["pos"][0], sc["rot"][0])  # arms the filter
    with pytest.raises(b200.B200Error):
        gm.depthPoseCallback(odd, sc["pos"][0], sc["rot"][0])  # the reference would read past the image end here


def test_half_voxel_origin_matches_the_oracle(b200, po, synth):
    """A map whose origin / resolution has fractional part 1/2 (size 12.9 m at 0.1 m): consecutive lattice cells of a ray
    can then fall into the same voxel, and the reference stops such a ray at its own flag."""
    size, res, gh = (12.9, 12.9, 3.05), 0.1, -0.05
    origin = (-6.45, -6.45, gh)
    rng_ = (5.5, 5.5, 4.5)
    gm = b200.GridMap(size, res, origin=origin, local_update_range=rng_, ground_height=gh)
    om = po.OracleMap(origin, size, res, local_update_range=rng_, ground_height=gh)
    intr = dict(fx=96.8, fy=96.8, cx=80.3, cy=60.9, use_depth_filter=0)
    gm.set_depth_params(**intr)
    om.set_depth_params(**intr)
    boxes = synth.pillar_boxes(8, origin, size, 30)
    r = np.random.default_rng(4)
    for f in range(6):
        pos = (-2.0 + 0.1 * f + r.uniform(-0.02, 0.02), 1.0, 1.0 + 0.05 * f)
        R, _ = synth.look_at(0.3 + 0.05 * f, 0.2)
        img = synth.render_depth(boxes, gh, pos, R, 120, 160, intr["fx"], intr["fy"], intr["cx"], intr["cy"], max_depth=8.0, dropout=0.03, seed=f)
        assert gm.depthPoseCallback(img, pos, R) == om.update_depth(img, pos, R)
        so, sg = om.depth_stats(), gm.depth_stats()
        assert (sg["projected"], sg["rays_cast"], sg["ray_steps"], sg["updates"]) == tuple(int(x) for x in so[:4]), f"frame {f}"
        assert np.array_equal(gm.occupancy(), om.occupancy()), f"frame {f}"
        assert np.array_equal(gm.inflate(), om.inflate()), f"frame {f}"
