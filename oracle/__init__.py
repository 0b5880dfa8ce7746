"""CPU oracle for the radiomic-feature hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product package
(``multimodal-isic_b200``) never imports it and has no CPU fallback.

PARITY UNPINNED: the arithmetic of the reference path lives in the third-party
package pyradiomics 3.1.0 (``/root/reference/params.yml:24``; imported at
``/root/reference/RadiomicExtractor.py:8``), which is neither vendored under
``/root/reference`` nor installable in this image.  This oracle restates its
published algorithm and is pinned only by the docstring
matrix examples of pyradiomics (tests/golden/) and by two independent
implementations of every matrix builder (NumPy here, plain C in
``cmatrices_oracle.c``) that must agree bit-exactly.
"""
