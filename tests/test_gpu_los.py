"""-m gpu: fastLOS booleans (and sample counts) bit-exact vs the oracle."""
import numpy as np
import pytest

from util import make_pair

pytestmark = pytest.mark.gpu


def _segments(c, n, seed, out_frac=0.1):
    rng = np.random.default_rng(seed)
    o, s = np.array(c["origin"]), np.array(c["size"])
    a = o + 0.2 + rng.random((n, 3)) * (s - 0.4)
    b = o + 0.2 + rng.random((n, 3)) * (s - 0.4)
    k = int(n * out_frac)
    b[:k] = o - 2.0 + rng.random((k, 3)) * (s + 4.0)          # some ends outside the map
    a[k:2 * k] = o - 1.0 + rng.random((k, 3)) * (s + 2.0)     # some starts outside
    short = slice(2 * k, 4 * k)
    b[short] = a[short] + rng.normal(0, 0.3, (2 * k, 3))      # short / grazing segments
    b[4 * k:4 * k + 10] = a[4 * k:4 * k + 10]                 # zero-length
    return a, b


@pytest.mark.parametrize("name", ["C3", "C1"])
def test_los_batch_matches_oracle(b200, po, synth, name):
    gm, om, c = make_pair(b200, po, synth, name)
    a, b = _segments(c, 200000, 17)
    vis_g, ns_g = gm.los_batch(a, b, return_samples=True)
    vis_o, ns_o = om.los_batch(a, b, nthreads=8, return_samples=True)
    assert 0.05 < vis_o.mean() < 0.95
    assert np.array_equal(vis_g, vis_o)
    assert np.array_equal(ns_g.astype(np.int64), ns_o)


def test_los_lattice_endpoints(b200, po, synth):
    # node-like endpoints: accumulated multiples of the planner resolution from an unsnapped start
    gm, om, c = make_pair(b200, po, synth, "C3")
    rng = np.random.default_rng(4)
    base = np.array([0.03, -0.07, 1.01])
    ia = rng.integers(-100, 100, (100000, 3)); ia[:, 2] = rng.integers(0, 40, 100000)
    ib = rng.integers(-100, 100, (100000, 3)); ib[:, 2] = rng.integers(0, 40, 100000)
    a, b = base + ia * 0.1, base + ib * 0.1
    assert np.array_equal(gm.los_batch(a, b), om.los_batch(a, b, nthreads=8))


def test_dk_sequence_is_not_k_times_step(po):
    dk = po.los_dk(1200)
    assert dk[0] == 0.0 and dk[1] == 0.05
    assert (dk != np.arange(1200) * 0.05).sum() > 1000   # repeated addition drifts


def test_dda_line_of_sight_matches_oracle(b200, po, synth):
    """b200_los_batch_dda against the oracle's RayCaster restatement: booleans and voxel counts."""
    gm, om, c = make_pair(b200, po, synth, "C3")
    r = np.random.default_rng(11)
    o, s = np.array(c["origin"]), np.array(c["size"])
    a = o - 0.2 + r.random((20000, 3)) * (s + 0.4)
    b = a + r.normal(0, 2.0, a.shape)
    b[:2000] = a[:2000]                                  # degenerate
    b[2000:4000, 1:] = a[2000:4000, 1:]                  # axis-aligned
    a[4000:6000] = np.round(a[4000:6000] / 0.1) * 0.1    # on voxel boundaries
    vg, ng = gm.los_batch_dda(a, b, return_cells=True)
    vo, no = om.los_batch_dda(a, b, return_cells=True)
    assert np.array_equal(vg, vo) and np.array_equal(ng, no)
    assert 0.05 < vg.mean() < 0.95
    # fastLOS and the traversal agree on most segments, not all (a sampled march can step over a corner voxel)
    vf = gm.los_batch(a, b)
    agree = (vf == vg).mean()
    assert agree > 0.9, agree
