"""oracle/ — TEST INFRASTRUCTURE ONLY (checkers, fixture generators, CPU baselines).

Nothing under this package is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
baseline legs of ``bench.py`` may import it.  The product (``corto_b200``) never does, and fails
loudly when its CUDA library is missing instead of falling back to anything here.

  crt_oracle.c / liboracle.so   plain-C restatement of the reference decode path (kind "port")
  ref_shim.cpp / _ref/*.so      the unmodified reference compiled in place        (kind "reference")
  refshim.py, pyoracle.py       ctypes bindings for the two
  meshgen.py                    deterministic synthetic meshes / clouds (SURVEY §8d generator)
"""
