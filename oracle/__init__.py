"""CPU oracle for the spectral path-tracing kernel -- TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
See oracle/oracle.cpp for the restatement itself and its parity-pin status ("parity unpinned by reference tests").
"""
