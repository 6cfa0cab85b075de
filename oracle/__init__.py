"""TEST INFRASTRUCTURE ONLY.

`oracle/` holds the CPU checkers for the hot path: `ref.py` binds the UNMODIFIED reference
compiled from /root/reference (oracle/_ref/libspade_ref.so) and `port.py` binds the plain-C
restatement (oracle/spade_oracle.c). Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package. The product (spade_b200/) never does.
"""
