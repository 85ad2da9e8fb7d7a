"""CPU oracle for the quantized Wan2.1 DiT hot path.  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Never imported by the product package."""
