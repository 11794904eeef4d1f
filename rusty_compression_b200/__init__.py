"""rusty_compression_b200 -- B200-native (sm_100a) implementation of rusty-compression's
randomized low-rank hot path.

* ``csrc/``   hand-written CUDA kernels + the C ABI declared in ``include/rc_api.h``
* ``_lib.py`` ctypes binding of that ABI (fails loudly if the library is not built)
* ``api.py``  host-side mirror of the reference's public surface (numpy in / numpy out)
* ``build.py`` in-tree nvcc build of ``librc_b200.so``

Importing the package does not touch the GPU; ``rusty_compression_b200.api`` does on first use.
"""
__all__ = ["api", "build", "_lib"]
