"""CPU oracle for the PRISMA band hot path (TEST INFRASTRUCTURE ONLY).

Everything under ``oracle/`` is a CPU restatement of the reference's algorithm
(patriciogonzalezvivo/prisma @ e00192dd) written as plain torch/numpy functional
code.  It exists to *check* the CUDA product in ``prisma_b200/``:

* only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
  ``--impl reference`` legs may import it;
* the product path (``prisma_b200/``, ``bands/``) never imports it and fails
  loudly when the CUDA library is missing.

Pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against the *reference modules themselves*, imported from
``/root/reference`` in the authoring container by
``oracle/tools/make_golden.py``, which also writes the small fixtures under
``tests/golden/`` that travel to the GPU box.
"""
