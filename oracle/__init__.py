"""TEST INFRASTRUCTURE — the CPU oracle for FlatNav's search hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this package, and only as the checker or the timed CPU baseline.  The product
package `flatnav_b200` never imports it and has no CPU fallback.

  oracle.port    ctypes wrapper over liboracle.so (flatnav_oracle.cpp): the CPU restatement
  oracle.refbin  runs the UNMODIFIED reference compiled into oracle/_ref/ (see Makefile)
"""
