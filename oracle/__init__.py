"""TEST INFRASTRUCTURE ONLY.

CPU restatement (plain PyTorch fp32 / numpy float64) of the hot path that the
CUDA library in ``graph_augmented_vision_transformers_b200`` accelerates.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package; the product
path never does (it raises if the CUDA library is missing).

Pinning status
--------------
* Attention / Block / VisionTransformer (``vit_oracle.py``): pinned against the
  reference's own modules (``/root/reference/src/models/vit.py``) run in the
  build container; committed fixtures in ``tests/golden/`` were produced by
  ``oracle/make_golden.py`` from the *imported reference*, not from this
  restatement.
* Graph construction / aggregation (``graph_oracle.py``): **parity unpinned** -
  the reference repository contains no graph code at all,
  so the frozen specification of SURVEY.md section 9 is the only oracle; ``knn_strict.c`` (plain C, built by
  ``knn_strict.py`` with gcc) fixes its fp32 accumulation order for the bit-exact index checks.
"""

# 1: SURVEY.md section 9 as written (fp32 accumulation order left to the GEMM library).
# 2: + the fp32 accumulation order of G1-G3 is FIXED (oracle/knn_strict.c: sequential FMA chains, normalise first, IEEE
#    division / sqrt) so that "neighbour indices bit-exact in fp32" is checkable on every row, not only on rows with a margin.
GRAPH_SPEC_VERSION = 2  # bump => every golden under tests/golden/graph_* is stale
