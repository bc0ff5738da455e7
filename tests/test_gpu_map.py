"""-m gpu: cloud -> inflated grid parity (bit-exact occupancy_buffer_inflate_) through the C-ABI."""
import numpy as np
import pytest

from util import make_pair

pytestmark = pytest.mark.gpu


def test_cloud_update_matches_oracle_c3(b200, po, synth):
    gm, om, c = make_pair(b200, po, synth, "C3")
    g, o = gm.inflate(), om.inflate()
    assert g.shape == o.shape == (256, 256, 64)
    assert o.sum() > 0
    assert np.array_equal(g, o)
    assert gm.local_bounds() == om.local_bounds()


@pytest.mark.parametrize("step", [12, 16, 32])
def test_point_strides(b200, po, synth, step):
    cloud = synth.uniform_cloud(11, (-12.8, -12.8, -0.01), (25.6, 25.6, 6.4), 50000, margin=0.5)
    gm, om, _ = make_pair(b200, po, synth, "C3", cloud=cloud, point_step=step)
    assert np.array_equal(gm.inflate(), om.inflate())
    assert gm.local_bounds() == om.local_bounds()


def test_lattice_aligned_points_position_space_inflation(b200, po, synth):
    # points whose coordinates sit exactly on voxel boundaries are where
    # "voxelise then dilate" differs from the reference's position-space inflation.
    rng = np.random.default_rng(3)
    ijk = rng.integers(-100, 100, size=(20000, 3))
    cloud = (ijk * 0.1).astype(np.float32)
    cloud[:, 2] = np.abs(cloud[:, 2]) % 6.0
    gm, om, _ = make_pair(b200, po, synth, "C3", cloud=cloud)
    assert np.array_equal(gm.inflate(), om.inflate())


@pytest.mark.parametrize("inflation", [0.0, 0.099, 0.25, 0.35])
def test_inflation_steps(b200, po, synth, inflation):
    cloud = synth.uniform_cloud(5, (-12.8, -12.8, -0.01), (25.6, 25.6, 6.4), 20000, margin=0.3)
    gm, om, _ = make_pair(b200, po, synth, "C3", cloud=cloud, inflation=inflation)
    assert np.array_equal(gm.inflate(), om.inflate())
    assert gm.local_bounds() == om.local_bounds()


def test_local_update_range_clears_only_the_box_and_filters_points(b200, po, synth):
    c = synth.CONFIGS["C3"]
    cloud = synth.config_cloud("C3")
    gm, om, _ = make_pair(b200, po, synth, "C3", cloud=cloud)
    # second frame: small range around a moved camera; everything outside the box must survive
    cloud2 = synth.uniform_cloud(9, c["origin"], c["size"], 30000)
    blob2 = synth.pack_cloud(cloud2)
    for cam, rng in [((2.0, -3.0, 1.0), (3.0, 2.5, 1.5)), ((-11.0, 11.0, 0.5), (4.0, 4.0, 4.0))]:
        gm.params.local_update_range[:] = rng  # not re-read: ranges are fixed at create time, so rebuild maps
        g2 = b200.GridMap(c["size"], c["resolution"], origin=c["origin"], local_update_range=rng, ground_height=c["origin"][2])
        o2 = po.OracleMap(c["origin"], c["size"], c["resolution"], local_update_range=rng, ground_height=c["origin"][2])
        g2.set_inflate(om.inflate())
        o2.set_inflate(om.inflate())
        g2.cloudCallback(blob2, cam)
        o2.update_cloud(blob2, cam)
        assert np.array_equal(g2.inflate(), o2.inflate())
        assert g2.local_bounds() == o2.local_bounds()


def test_empty_cloud_and_nan_camera_are_noops(b200, po, synth):
    gm, om, _ = make_pair(b200, po, synth, "C3")
    before = gm.inflate()
    gm.cloudCallback(np.zeros((0, 4), np.float32), (0, 0, 1))
    gm.cloudCallback(synth.pack_cloud(synth.config_cloud("C3")), (float("nan"), 0, 1))
    assert np.array_equal(gm.inflate(), before)


def test_nan_inf_and_out_of_map_points(b200, po, synth):
    cloud = synth.uniform_cloud(2, (-12.8, -12.8, -0.01), (25.6, 25.6, 6.4), 5000, margin=3.0)
    cloud[::7, 0] = np.nan
    cloud[::11, 1] = np.inf
    cloud[::13, 2] = -np.inf
    gm, om, _ = make_pair(b200, po, synth, "C3", cloud=cloud)
    assert np.array_equal(gm.inflate(), om.inflate())
    assert gm.local_bounds() == om.local_bounds()


def test_clear_box_and_reset(b200, po, synth):
    gm, om, _ = make_pair(b200, po, synth, "C3")
    for mn, mx in [((-3.05, -2.0, 0.33), (4.0, 1.0, 2.71)), ((-100, -100, -100), (0, 0, 0)), ((5, 5, 5), (4, 4, 4))]:
        gm.resetBuffer(mn, mx)
        om.clear_box(mn, mx)
        assert np.array_equal(gm.inflate(), om.inflate())
    gm.resetBuffer()
    assert gm.inflate().sum() == 0
    assert gm.local_bounds() == ((0, 0, 0), (255, 255, 63))


def test_occupancy_inmap_index_reads(b200, po, synth):
    gm, om, c = make_pair(b200, po, synth, "C3")
    rng = np.random.default_rng(0)
    o, s = np.array(c["origin"]), np.array(c["size"])
    pos = o - 0.5 + rng.random((200000, 3)) * (s + 1.0)
    # boundary cases of isInMap: exactly min+1e-4 / max-1e-4 and their neighbours
    edge = np.array([[o[0] + 1e-4, 0, 1], [np.nextafter(o[0] + 1e-4, -1), 0, 1], [o[0] + s[0] - 1e-4, 0, 1],
                     [np.nextafter(o[0] + s[0] - 1e-4, 1e9), 0, 1], [0.6000000000000001, 0.30000000000000004, 1.0]])
    pos = np.concatenate([pos, edge])
    assert np.array_equal(gm.getInflateOccupancy(pos), om.get_inflate_occupancy(pos))
    assert np.array_equal(gm.isInMap(pos), om.is_in_map(pos))
    assert np.array_equal(gm.posToIndex(pos), om.pos_to_index(pos))


def test_upload_download_roundtrip_odd_dims(b200, po, synth):
    # Nz = 50 (not a multiple of 32) as in C1
    c = synth.CONFIGS["C1"]
    gm = b200.GridMap(c["size"], c["resolution"], origin=c["origin"])
    assert gm.dims == (250, 250, 50)
    rng = np.random.default_rng(1)
    grid = (rng.random(gm.dims) < 0.1).astype(np.uint8)
    gm.set_inflate(grid)
    assert np.array_equal(gm.inflate(), grid)


def test_c2_full_size_cloud_1m_points(b200, po, synth):
    cloud = synth.config_cloud("C2", n_points=1_000_000)
    assert cloud.shape == (1_000_000, 3)
    gm, om, _ = make_pair(b200, po, synth, "C2", cloud=cloud, rng=(60.0, 60.0, 20.0))
    g, o = gm.inflate(), om.inflate()
    assert g.shape == (512, 512, 128)
    assert np.array_equal(g, o)
    assert gm.local_bounds() == om.local_bounds()
    # idempotence: the same frame again gives the same grid (clear + set-union)
    gm.cloudCallback(synth.pack_cloud(cloud), (0.0, 0.0, 1.0))
    assert np.array_equal(gm.inflate(), o)


def test_publish_map_inflate_matches_oracle(b200, po, synth):
    """publishMapInflate (grid_map.cpp:851-900): ordered export of the occupied voxels; the oracle's export is pinned
    against the reference's own publisher in tests/test_oracle_vs_ref.py."""
    c = synth.CONFIGS["C3"]
    for rng, cam in [((1e3, 1e3, 1e3), (0.0, 0.0, 1.0)), ((4.0, 3.0, 1.5), (2.0, -3.0, 1.0))]:
        gm, om, _ = make_pair(b200, po, synth, "C3", rng=rng, cam=cam)
        for all_info in (False, True):
            for trunc in (2.4, 0.77, -1.0):
                want, n = om.export_inflate_cloud(1 if all_info else 0, trunc)
                got = gm.publishMapInflate(all_info, 1, trunc)
                assert len(got) == n
                assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        assert len(gm.publishMapInflate(cap=10)) == 10      # truncated buffer: first 10 points, still ordered


def test_set_occupied_and_set_occupancy_match_oracle(b200, po):
    """GridMap::setOccupied / setOccupancy (grid_map.h:310-337), batched; the oracle is pinned to the reference's own
    methods in test_oracle_vs_ref.py (same position generator)."""
    from util import set_positions as _set_positions
    size, res, gh = (6.4, 4.8, 3.2), 0.1, -0.01
    origin = (-size[0] / 2, -size[1] / 2, gh)
    gm = b200.GridMap(size, res, origin=origin, ground_height=gh)
    om = po.OracleMap(origin, size, res, 0.099, (1e3, 1e3, 1e3), gh)
    gm.set_depth_params()
    om.set_depth_params()
    pos, occ = _set_positions(5, origin, size, 4000)
    gm.setOccupied(pos[:2000])
    om.set_occupied(pos[:2000])
    assert np.array_equal(gm.inflate(), om.inflate())
    gm.setOccupancy(pos, occ)
    om.set_occupancy(pos, occ)
    assert np.array_equal(gm.occupancy(), om.occupancy())
    assert np.array_equal(gm.getOccupancy(pos), om.get_occupancy(pos))
    # a later depth frame still works on the same state arrays (the per-voxel scratch was left clean)
    gm.setOccupancy(pos[:10], 1.0)
    om.set_occupancy(pos[:10], 1.0)
    assert np.array_equal(gm.occupancy(), om.occupancy())
