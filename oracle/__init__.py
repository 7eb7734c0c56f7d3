"""CPU restatement of the reference's hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; the product (motifs.jl_b200/) never does.  PARITY UNPINNED: the reference ships no
tests or golden vectors and cannot run here (no Julia), so every function restates the reference
source it cites instead of being checked against reference output.
"""
