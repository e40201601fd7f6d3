"""TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's hot path (the parity checker).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this package.  The product package (liana_py_b200) never does.
"""
