"""CPU restatements of the reference's hot path -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Import policy: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only.
"""
