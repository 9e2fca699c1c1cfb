"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference algorithms on the scene-to-model dense correspondence path
(SURVEY.md section 8), each function citing the reference file:line it follows, plus loaders for the
reference's own native kNN compiled from /root/reference into oracle/_ref/ (never copied).

Import rule (DESIGN.md "Oracle"): only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package.  The product package
(geometric-aware-dense-matching_b200/, importable as gadm_b200) never does and has no CPU fallback.

Pinning status (the reference ships no tests, golden vectors or fixtures for this path):
  * kNN 3-D   : pinned against the reference's own compiled knn_.cxx + nanoflann (oracle/_ref).
  * matching  : the arithmetic is 5 lines of stock torch (evaluator.py:89-93); restated with the same
                torch calls; the reference module itself is not importable (mmcv/detectron2 absent).
  * dgcnn kNN : pinned against models/dgcnn.py imported from /root/reference (fixtures in tests/golden).
  * pointops  : "parity unpinned" -- CUDA sources are absent upstream; only the pure-torch
                KNNQueryNaive body and the Grouping docstring define the semantics.
  * soft correspondence (weights, soft_xyz): an extension defined here, not in the reference.
"""
