"""oracle/ - CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import anything from here.  The product package (jukebox_b200/) never does.
"""
