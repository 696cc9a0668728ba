"""Host-side scene plumbing of the streamed step (no GPU): shape signatures, in-place refill, cache invalidation, aliasing."""
import numpy as np
import pytest
import torch

from gasfm_b200.scene import Scene
from gasfm_b200 import synthetic


def _scene(m, n, n_obs, seed):
    idx, vals = synthetic.synthetic_observations(m, n, n_obs, seed)
    return Scene.from_observations(idx, vals, m, n), idx, vals


def test_signature_and_copy_from_refill_in_place():
    a, idx_a, vals_a = _scene(12, 400, 3000, seed=1)
    cols = 399 - idx_a[1]
    order = np.lexsort((cols, idx_a[0]))
    b = Scene.from_observations(np.stack((idx_a[0][order], cols[order])), vals_a[order] * 0.5, 12, 400)
    assert a.signature() == b.signature()
    ptrs = {k: v.data_ptr() for k, v in a._buffers().items()}
    a.copy_from(b, non_blocking=False)
    assert {k: v.data_ptr() for k, v in a._buffers().items()} == ptrs          # same storage, new contents
    assert torch.equal(a.x.indices, b.x.indices) and torch.equal(a.x.values, b.x.values)
    assert torch.equal(a.graph_wrappers["scenepoint2global"].valid_indices, b.graph_wrappers["scenepoint2global"].valid_indices)
    # the observation graphs alias x.indices: refilled with it, not separately
    assert a.graph_wrappers["proj2view"].valid_indices is a.x.indices
    c, _, _ = _scene(12, 400, 2500, seed=2)
    assert c.signature() != a.signature()
    with pytest.raises(ValueError, match="differ in shape"):
        a.copy_from(c)


def test_to_keeps_the_aliases_and_invalidate_drops_cached_structure():
    a, _, _ = _scene(10, 300, 2000, seed=3)
    moved = a.to("cpu")
    for k in ("proj2view", "proj2scenepoint"):
        assert moved.graph_wrappers[k].valid_indices is moved.x.indices        # one copy of the [2,E] indices, not three
    from gasfm_b200.index import _INDEX_ATTR
    from gasfm_b200.models.layers import _PLAN_ATTR
    setattr(moved.x, _INDEX_ATTR, object())
    setattr(moved.graph_wrappers["view2global"], _PLAN_ATTR, object())
    moved.invalidate()
    assert not hasattr(moved.x, _INDEX_ATTR) and not hasattr(moved.graph_wrappers["view2global"], _PLAN_ATTR)


def test_deferred_index_validation_bookkeeping():
    """The streamed step builds the observation index inside a captured graph, so the status word cannot be read back during
    the build: the context collects what to check later; the status bits map to the same errors as the immediate check."""
    from gasfm_b200 import index
    assert index._DeferredValidation.active is None
    with index.deferred_validation() as outer:
        assert index._DeferredValidation.active is outer
        with index.deferred_validation() as inner:
            assert index._DeferredValidation.active is inner
        assert index._DeferredValidation.active is outer and outer.entries == []
    assert index._DeferredValidation.active is None
    index.raise_for_status(0, 3, 4)
    with pytest.raises(ValueError, match="out of range"):
        index.raise_for_status(1, 3, 4)
    with pytest.raises(ValueError, match="row-major sorted"):
        index.raise_for_status(2, 3, 4)
