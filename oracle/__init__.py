"""CPU oracle for the CTR embedding hot path — TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy, fp32 with fp64 shadows where noted) of
the reference's algorithm for the path BASELINE.json names.  It is the checker.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it.  Nothing under ``recommender_b200/``
imports it; the product path fails loudly when the CUDA library is missing.

PARITY UNPINNED (partially): the reference ships no tests or golden vectors and
its arithmetic lives in the un-vendored ``tensorflow~=2.2.0`` wheel
(ctr/requirements.txt:1), which is absent here.  What *is* pinned:
``ctr/model.py`` and ``ctr/layers.py`` are imported byte-for-byte under a tiny
``tensorflow`` shim (oracle/tf_shim) and their forward outputs + autograd
gradients are frozen as fixtures in tests/golden/ (tests/golden/make_golden.py).
The TF-internal pieces (IndexedSlices dedup, Keras Adam/Adagrad, BCE form,
SURVEY Appendix A) are restated from the published Keras semantics; the Adam
formula and the first-occurrence dedup order are additionally cross-checked
against two independent implementations of the same definitions that ARE in
this image (scikit-learn's AdamOptimizer, pandas.factorize:
tests/test_oracle_golden.py) — known-answer checks, not the reference itself.

PINNED: the input side.  oracle/criteo_oracle.py reproduces bit for bit what
the reference's own ``build_vocab`` / ``write_tfrecord`` (ctr/tfrecord_io.py,
imported byte-for-byte under the shim) produce on synthetic Criteo files
(tests/golden/criteo_tsv.npz, tests/golden/make_golden_criteo.py).
"""
