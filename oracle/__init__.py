"""CPU oracle for the GPnet dense-GP hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product path: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / CPU comparator.  The product (``gpnet_b200``)
never imports this package and fails loudly when its CUDA library is missing.

Parity status: the reference (filippocastelli/GPnet) has NO tests and no golden vectors for the
graph hot path, so the oracle is pinned by (a) outputs of the *unmodified*
reference classes run in the build container under the three shims of ``oracle/ref_shim.py``
(fixtures in ``tests/golden/*.npz``, generator ``oracle/make_golden.py``), (b) the reference's only
known-answer value, ``report_notebook.ipynb`` cell 49 (``-77.22939013``), and (c) independent
cross-checks (spectral ``eigh`` restatement, central finite differences of the gradient).
"""
