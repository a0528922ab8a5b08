"""CPU parity oracle for the B200 Stratego engine -- TEST INFRASTRUCTURE ONLY.

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  Nothing under stratego_env_b200/ imports it.
"""
