"""CPU oracle for the multi-view reconstruction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``skiing_analysis_pytorch_b200/`` may import
this package: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and there only as the checker or as the
CPU arm that is timed *beside* the CUDA path, never as the thing shipped.

The oracle is an fp64 numpy restatement of the arithmetic the reference delegates to
``cv2.triangulatePoints`` / ``cv2.projectPoints`` / ``cv2.Rodrigues`` / ``np.linalg.svd`` /
torch broadcasting, each function citing the reference file:line it follows.

Parity pinning: the reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF:
``oracle/make_golden.py`` imports the reference's own functions from ``/root/reference``
(authoring container only) and freezes their outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every restatement here against those vectors.
The Levenberg-Marquardt solver has NO reference implementation (``run_local_ba`` is an
undefined symbol, vggt/multi_view_process.py:553) - its *cost* is pinned by the
reference's ``reprojection_loss`` (bundle_adjustment/loss.py:90-94), its *trajectory* is
"parity unpinned" and defined by ``oracle/lm.py`` (spec in DESIGN.md section 5).
"""
