"""CPU oracle for the GASFM graph-attention hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker or as the timed
CPU baseline.  ``gasfm_b200`` never imports this package.

Parity status
-------------
* Everything that lives in the reference tree (index build, graph wrappers,
  layer wiring, output decoding) is PINNED: ``tests/golden/*.npz`` were produced
  by importing the unmodified reference sources from ``/root/reference/code``
  (``tests/golden/make_golden.py``) and the oracle reproduces them.
* The per-edge arithmetic of ``torch_geometric.nn.GATv2Conv`` lives in a
  third-party dependency that is absent from ``/root/reference`` and from this
  image (conda ``pyg::pyg``, unpinned in ``environment.yml:50``; the code
  mentions PyG 2.2.0 at ``code/train.py:236``).  It is restated from PyG's
  published algorithm in ``oracle/gatv2conv.py`` and anchored on the reference's
  call sites (``code/models/layers.py:304,329,401,426,506,521,550,566``).  That
  one piece is "parity unpinned" against a real PyG build.
"""
