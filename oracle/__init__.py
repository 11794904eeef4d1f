"""CPU oracle for the rusty-compression randomized low-rank hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker (or the timed
CPU baseline).  The product path (``rusty_compression_b200``) never imports it
and fails loudly when its CUDA library is missing.

What it is
----------
A numpy/scipy restatement of the reference crate's control flow
(``/root/reference/src/*.rs``, cited function by function in
``oracle/reference_path.py``) that calls the *same LAPACK routines the crate
calls* -- ``?geqp3``, ``?orgqr/?ungqr``, ``?gesdd`` (jobz='S'), ``?trtrs`` --
through ``scipy.linalg.lapack`` (bundled OpenBLAS).  The reference itself is
Rust and cannot be compiled here (no cargo/rustc, dependencies un-vendored and
un-pinned: ndarray 0.15.*, ndarray-linalg 0.16.*, lax 0.*, lapack 0.*,
rand 0.8, rand_distr 0.4 -- Cargo.toml:17-26, no Cargo.lock).

Pinning status
--------------
* ``permutation``: pinned bit-exactly by the reference's own known-answer
  tests (src/permutation.rs:192-239), carried over verbatim in
  ``tests/test_oracle_permutation.py``.
* ``philox``: pinned by the published Random123 Philox4x32-10 known-answer
  vectors (``tests/test_oracle_philox.py``).
* Everything numerical (pivoted QR, SVD, IDs, samplers): the reference holds
  NO golden vectors (every test draws from ``thread_rng()`` and asserts
  properties; random_sampling.rs has no tests at all), so for those functions
  **parity is unpinned** beyond (i) the reference's 89 property tests, ported
  in ``tests/test_oracle_reference_properties.py`` with the reference's shapes
  and thresholds, and (ii) the one published result, the adaptive-sampling
  convergence curve (rank ~115 at tol 1e-5 on the 500x200 example,
  residuals.png), reproduced in ``tests/test_oracle_adaptive_example.py``.
"""
