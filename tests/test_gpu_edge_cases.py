"""Edge cases of the explicit path against the oracle: the smallest meshes the reference's builders accept (2 to 5 nodes per axis: zero,
one or a few assignable cells, every cell next to a boundary), ranges that the stencil's prepare() shrinks to nothing, one-cell-thick
2-D / 3-D fields (register-window and TMA skeletons must hand over to the direct skeleton) and reductions over empty / one-cell ranges.
The reference walks such ranges with rangeFor over an empty or tiny DS::Range (RangeFor.hpp:29-84): nothing is written outside it."""
import numpy as np
import pytest

from opflow_b200 import capi, host
from opflow_b200.host import D1FirstOrderCentered, D2SecondOrderCentered, d2x, d2y, d2z, dx
from helpers import assert_same, dirc, gpu_storage, make_pair, set_both

pytestmark = pytest.mark.gpu
MODES = [(capi.MODE_EXACT, True), (capi.MODE_FAST, False)]


def lap(g, dim):
    D = D2SecondOrderCentered
    e = d2x(D, g)
    if dim >= 2:
        e = e + d2y(D, g)
    if dim >= 3:
        e = e + d2z(D, g)
    return e


@pytest.mark.parametrize("mode,exact", MODES)
@pytest.mark.parametrize("dims", [[2], [3], [4], [5], [2, 2], [3, 3], [2, 5], [5, 2], [3, 4], [2, 2, 2], [3, 3, 3], [3, 2, 5], [4, 3, 2], [70, 2, 3], [3, 70, 2]])
def test_ftcs_on_tiny_meshes(engine, oracle, mode, exact, dims):
    """aliased FTCS step with Dirichlet walls on meshes of 2..5 nodes per axis (and one long axis beside two degenerate ones)"""
    host.set_mode(mode)
    dim = len(dims)
    g, o = make_pair(dims, [0] * dim, [1] * dim, bc=dirc(dim))
    set_both(g, o)
    g.updatePadding(), o.update_padding()
    e = g + 0.01 * lap(g, dim)
    for _ in range(3):
        g.assign(e)
        oracle.assign(o, e.signature(), [o] * (dim + 1), [0.01])
    a, b = gpu_storage(g, o)
    assert_same(a, b, exact, what=f"tiny mesh {dims}")


@pytest.mark.parametrize("mode,exact", MODES)
@pytest.mark.parametrize("dims,loc", [([3, 3], [1, 1]), ([2, 4], [1, 1]), ([4, 2, 3], [1, 1, 1]), ([3, 3], [0, 1]), ([2, 2, 2], [1, 0, 1])])
def test_cell_centred_tiny_meshes_with_ghosts(engine, oracle, mode, exact, dims, loc):
    """Center / mixed staggering with ext 1 on 1..3 cells per axis: every stencil tap of every cell is a ghost or a wall value"""
    host.set_mode(mode)
    dim = len(dims)
    bc = {(d, s): (host.BCType.Neum if d == 0 else host.BCType.Dirc, 0.25 * (d + 1) * (1 - 2 * s)) for d in range(dim) for s in range(2)}
    g, o = make_pair(dims, [0] * dim, [1] * dim, loc=loc, bc=bc, ext=1)
    set_both(g, o)
    g.updatePadding(), o.update_padding()
    e = g + 0.01 * lap(g, dim)
    for _ in range(2):
        g.assign(e)
        oracle.assign(o, e.signature(), [o] * (dim + 1), [0.01])
    a, b = gpu_storage(g, o)
    assert_same(a, b, exact, what=f"tiny centred mesh {dims} loc {loc}")


def test_reductions_over_tiny_and_empty_ranges(engine):
    """rangeReduce over one cell, and over an empty range: the identity of the reduction (RangeFor.hpp:99-141 starts from the caller's
    identity element; sum 0, max -inf / min +inf are what the reference's own callers pass)"""
    host.set_mode(capi.MODE_EXACT)
    g, _ = make_pair([5, 4], [0, 0], [1, 1], bc=dirc(2))
    vals = np.asfortranarray(np.arange(20, dtype=np.float64).reshape(5, 4, order="F") - 7.0)
    g.from_numpy(vals)
    one = capi.Range.make([2, 1], [3, 2])
    assert host.rangeReduce(g, capi.RED_SUM, one) == vals[2, 1]
    assert host.rangeReduce(g, capi.RED_MAX, one) == vals[2, 1]
    assert host.rangeReduce(g, capi.RED_ABSMAX, one) == abs(vals[2, 1])
    empty = capi.Range.make([2, 1], [2, 2])
    assert host.rangeReduce(g, capi.RED_SUM, empty) == 0.0
    assert host.rangeReduce(g, capi.RED_ABSMAX, empty) == 0.0
    whole = g.localRange  # the wall nodes hold the Dirichlet value after the upload's updatePadding, not what was uploaded
    now = g.to_numpy(whole)
    assert now[0, 0] == 1.0 and now[2, 1] == vals[2, 1]
    assert host.rangeReduce(g, capi.RED_SUM, whole) == sum(now.reshape(-1, order="F").tolist()) or abs(host.rangeReduce(g, capi.RED_SUM, whole) - now.sum()) <= 1e-12
    assert host.rangeReduce(g, capi.RED_MIN, whole) == now.min()


def test_assignment_whose_prepared_range_is_empty_writes_nothing(engine):
    """a stencil on a 2-node Dirichlet axis has no assignable cell: the call succeeds and leaves the field as it was; adding a Center field
    and its Corner-located derivative is a location mismatch, reported (OPF_ERR_LOC) like the reference's OP_ASSERT in prepare()"""
    host.set_mode(capi.MODE_EXACT)
    g, o = make_pair([2, 6], [0, 0], [1, 1], bc=dirc(2))
    set_both(g, o)
    before = g.to_numpy(g.getLocalReadableRange()).copy()
    g.assign(g + 0.5 * lap(g, 2))
    assert np.array_equal(g.to_numpy(g.getLocalReadableRange()), before)
    c, _ = make_pair([6, 6], [0, 0], [1, 1], loc=[1, 1], bc={(d, s): (host.BCType.Neum, 0.) for d in range(2) for s in range(2)}, ext=1)
    with pytest.raises(capi.EngineError):
        c.assign(c + dx(D1FirstOrderCentered, c))
