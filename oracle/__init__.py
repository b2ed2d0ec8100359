"""CPU oracle of the MULLS registration hot path — TEST INFRASTRUCTURE (see mulls_oracle.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
