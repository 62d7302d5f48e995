"""TEST INFRASTRUCTURE (see oracle/port/oracle.h).  Python access to the CPU oracle:

* ``oracle.port``  -- ctypes bindings of ``oracle/liboracle.so`` (our C restatement; travels to the GPU box)
* ``oracle.refshim`` -- ctypes bindings of ``oracle/_ref/libref_handle.so`` (the reference's own loaders,
  compiled unmodified; exists only where /root/reference was available at build time)

Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
