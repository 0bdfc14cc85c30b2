"""CPU oracle package -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, on the CPU, the reference's algorithm for the hot path (see each
module's header for the reference file:line it follows).  May be imported only
by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl
reference` legs.  seed_rl_b200/ never imports it.
"""
