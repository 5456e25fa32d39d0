"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's hot path, used as the parity checker by
tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of
bench.py.  Nothing in claude_semantic_search_b200/ may import this package.
"""
