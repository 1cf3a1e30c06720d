"""CPU oracle for the XNet batched implicit step.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (xnet_b200/) may import, link or
execute anything in this directory; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs do, and only as the checker / reported baseline.
"""
