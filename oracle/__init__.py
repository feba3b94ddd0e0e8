"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's integral-regression path used as the parity checker.
Nothing under the product package imports this; only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may.
"""
