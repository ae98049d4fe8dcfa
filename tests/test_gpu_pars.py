"""Parsimony on the B200 through the C ABI (plk_pars_*): bit-exact against the dumps of the unmodified reference's
src/pars.c (tests/golden/pars) and against the oracle on seeded synthetic inputs.  `-m gpu`."""
import numpy as np
import pytest

import pars_checks as pk
from oracle_backend import OracleBackend
from phyml_b200.engine import Engine, EngineError

pytestmark = pytest.mark.gpu


def make(case, **kw):
    c, g = pk.load(case)
    return c, g, Engine(c.n_otu, c.P, c.ns, c.ncatg, c.tree.n_clv_handles, c.tree.n_edges, **kw)


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("general", [False, True])
@pytest.mark.parametrize("case", pk.PARS_CASES)
def test_full_traversal_exact(case, general, fused):
    pk.check_full(*make(case), general, fused)


@pytest.mark.parametrize("general", [False, True])
def test_single_updates_exact(general):
    pk.check_single_updates(*make("nucleic_hky"), general)


@pytest.mark.parametrize("general", [False, True])
def test_sharded_one_process(general):
    """plk_create_sharded: three shards on device 0 own contiguous pattern blocks; totals are the all-shard sums"""
    pk.check_full(*make("proteic_lg", devices=[0, 0, 0]), general, True)


@pytest.mark.parametrize("n_otu,P,ns,general", [(30, 1, 4, False), (17, 333, 4, True), (40, 70001, 4, False),
                                                (12, 5000, 20, True), (25, 20011, 20, False), (8, 777, 7, True),
                                                (6, 400003, 4, False), (5, 700001, 4, False)])
def test_vs_oracle_synthetic(n_otu, P, ns, general):
    """ragged sizes, 1 pattern, generic state counts, the 2- and 4-patterns-per-thread variants of the Fitch kernel"""
    tree, ui, w, step = pk.random_case(n_otu, P, ns, seed=n_otu + P)
    args = (tree.n_otu, P, ns, 1, tree.n_clv_handles, tree.n_edges)
    a = pk.run_random(tree, ui, w, step, Engine(*args), general)
    b = pk.run_random(tree, ui, w, step, OracleBackend(*args), general)
    assert a[0] == b[0] and (a[1] == b[1]).all() and a[2] == b[2]
    c = pk.run_random(tree, ui, w, step, Engine(*args), general, split=True)
    assert c[0] == b[0] and (c[1] == b[1]).all()


@pytest.mark.parametrize("devices", [None, [0, 0]])
def test_fractional_weights_truncate_like_the_reference(devices):
    """`tree->c_pars += site_pars * wght` on an int (src/pars.c:46): the serial chain kernel"""
    tree, ui, w, step = pk.random_case(14, 4099, 4, seed=8, frac_weights=True)
    args = (tree.n_otu, 4099, 4, 1, tree.n_clv_handles, tree.n_edges)
    a = pk.run_random(tree, ui, w, step, Engine(*args, devices=devices), False)
    b = pk.run_random(tree, ui, w, step, OracleBackend(*args), False)
    assert a[0] == b[0] and (a[1] == b[1]).all() and a[2] == b[2]


def test_errors():
    c, g, eng = make("nucleic_hky")
    with pytest.raises(EngineError):
        eng.pars_update([(0, 1, 2)])  # plk_pars_create not called
    eng.pars_create(c.tree.n_clv_handles)
    with pytest.raises(EngineError):
        eng.pars_update([(0, 1, 2)])  # children never written
    with pytest.raises(EngineError):
        eng.pars_update([(c.tree.n_clv_handles, 1, 2)])  # handle out of range
    with pytest.raises(EngineError):
        eng.pars_set_buffer(1, p_pars=np.zeros((c.P, c.ns), dtype=np.int32))  # no step matrix given
    with pytest.raises(EngineError):
        eng.get_site_pars()
