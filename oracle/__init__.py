"""CPU oracle — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (input production through the reference writer,
the cpu_baseline leg and --impl reference) import this package.  The product package libzseek_b200
never does; see oracle/zsk_oracle.c and oracle/refdrive.c for what each library is.
"""
