"""CPU oracle for the PCLSegmentation inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker (or as the
timed CPU baseline), never as the thing shipped.  The product path
(``pclsegmentation_b200``) never imports this package and fails loudly when its
CUDA library is missing.

Parity pinning status (see DESIGN.md "Oracle"):

* ``oracle.projection``  - PINNED against the reference itself: the reference's
  ``LaserScan`` / ``SemLaserScan`` (dataset_convert/laserscan_semantic_kitti.py) is
  importable in the build container and the committed fixtures
  ``tests/golden/projection_*.npz`` were produced by running it
  (``tests/golden/make_projection_golden.py``).
* ``oracle.nn`` and ``oracle.confusion`` - PARITY UNPINNED: the arithmetic lives in
  tensorflow-gpu==2.9.1 (requirements.txt:1) which is not vendored and not
  installable here; the reference ships no tests, golden logits or weights.  The
  restatement follows the published TF/Keras 2.9 op semantics and is self-checked
  with hand-computed cases (tests/test_oracle_nn.py).
"""
