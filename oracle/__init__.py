"""CPU oracle for the EVREAL hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the arithmetic of the three hot stages of the
reference's ``eval.py`` loop (voxelizer -> recurrent reconstruction network ->
per-frame metrics) plus the small glue between them.  It exists so the CUDA
path in ``evreal_b200`` can be checked for parity on a machine that does not
have ``/root/reference``.

Rules (enforced by ``tests/test_layout.py``):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
    ``cpu_baseline`` / ``--impl reference`` legs may import this package;
  * nothing under ``evreal_b200/`` imports it; the product path has no CPU
    fallback and raises if the CUDA library is missing.

Pinning status (see DESIGN.md "Oracle"):
  * voxelizer, window tables, crop/pad, normalisation, all networks: pinned
    against the *imported* reference (``tools/make_golden.py`` ran the real
    ``/root/reference`` modules in the build container; vectors committed in
    ``tests/golden/``).
  * MSE / SSIM: the reference calls scikit-image (unpinned, absent offline).
    The restatement follows the published algorithm and is pinned against
    ``scipy.ndimage`` (which scikit-image itself calls) -- "parity unpinned"
    with respect to scikit-image proper.
  * LPIPS: pyiqa and its weights are absent offline -- "parity unpinned";
    architecture-level restatement with seeded weights only.
"""
