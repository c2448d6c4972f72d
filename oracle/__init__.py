"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements (numpy / torch-CPU) of the DCFP scoring -> mask -> gather path,
used as the *checker* for the CUDA product path in ``dcfp_b200``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from this package.  Nothing in
``dcfp_b200`` imports it: the product path has no CPU fallback and raises when the
CUDA library is missing.

Pinning status (details in DESIGN.md section "Oracle"):

* ``eic_ref``, ``mask_ref``, ``gather_ref``, ``scoring_ref``: the reference ships no
  tests / golden vectors, so these restatements are pinned against outputs of the
  *unmodified reference code executed in the build container* (``oracle/ref_compat.py``
  + ``tests/golden/make_golden.py``; fixtures are committed under ``tests/golden/``:
  ``eic_steps.npz``, ``prune_c{1..4}.npz``, ``prune_c1_beta.npz``, ``sweep_c{1..4}.npz``,
  ``scoring_small.npz``).
* ``class_stats_ref`` (``bwd`` value functor): pinned through the identity
  ``sum_k S1[k, c] == bn.weight.grad`` of the reference's autograd path.
* ``class_stats_ref`` (``fwd`` value functor): there is no reference counterpart
   -- **parity unpinned**; the spec is the builder's own.
"""
