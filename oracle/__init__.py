"""TEST INFRASTRUCTURE — CPU oracle for the SR-network hot path of yangsenwxy/VSR.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package, and only as the checker / reported baseline.  The product (vsr_b200/)
never imports it and has no CPU path.

Contents
  restated.py        plain-torch functional restatement of the reference nets, losses, metrics
                     (each function cites the reference file:line it follows)
  load_reference.py  stub loader that imports the *real* reference modules from /root/reference
                     (exists only in the build container, not on the GPU box)
  make_golden.py     runs the real reference on seeded inputs and writes tests/golden/*.pt

Parity pin: restated.py is checked against the golden vectors produced by the real reference
(tests/test_oracle.py) and, where /root/reference is mounted, against the live reference.
"""
