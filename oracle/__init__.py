"""ORACLE -- test infrastructure only.  See oracle/mplb_oracle.h for the scope statement.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (mplambda_b200/) never does.
"""
