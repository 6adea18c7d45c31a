"""CPU oracle for the Streets4MPI per-day traffic-assignment hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``streets4mpi_b200/`` may import this
package; it is imported by ``tests/``, by ``__graft_entry__.smoke()`` and by the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- always as the
checker or the reported CPU baseline, never as the product path.

Parity status (see DESIGN.md "Oracle"):

* K1 weights, the trip walk, ``merge_arrays`` and the road-adaptation sweep are
  PINNED: ``tests/golden/make_golden.py`` executes the reference's own source
  text (``/root/reference/simulation.py``, ``streetnetwork.py``, ``utils.py``,
  ``settings.py``) under Python 3 and stores what it returns; the restatement
  here is checked against those fixtures.
* The Dijkstra itself lives in python-graph (PyPI ``python-graph-core``, no
  version pinned by the reference, source absent from ``/root/reference``).
  ``oracle/pygraph_standin.py`` restates its published algorithm from memory:
  **parity unpinned** at that boundary (ties only; where the shortest path is
  unique every correct SSSP gives the same tree).
"""
