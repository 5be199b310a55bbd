"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the CNE+SUE hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker (or as the timed CPU baseline), never as
the implementation that is shipped.  The product (``nnr_b200``) fails loudly when its CUDA
library is missing; it never falls back to this code.

Contents
--------
``nnr_oracle.py``     plain-PyTorch (CPU, fp32/fp64) restatement of the reference's CNE news
                      encoder, SUE user encoder, dot-product click predictor, loss and
                      clip+Adam step, each function citing the reference file:line it follows.
``graph.py``          numpy restatement of the history-graph / cluster-index construction
                      (reference ``MIND_corpus.py:162-216``).
``reference_import.py`` imports the *unmodified* reference modules from ``/root/reference`` with
                      a ``torch_scatter`` shim and a stub config.  Works only where
                      ``/root/reference`` exists (the build container); used to validate the
                      restatement and to generate ``tests/golden/*``.

Pinning status: the reference ships no tests, golden vectors or fixtures, so
the restatement is pinned against outputs of the reference itself executed in the build container:
``tests/golden/make_golden.py`` runs the real ``model.Model`` and stores its logits / loss /
gradient digests, and ``tests/test_oracle.py`` checks the restatement against those files (and,
where ``/root/reference`` is present, against the live reference).  ``torch_scatter`` 2.0.9 is an
absent third-party dependency: its two call sites are restated from the package's published
semantics (see ``nnr_oracle.scatter_softmax`` / ``scatter_sum``) -- that part is "parity unpinned".
"""
