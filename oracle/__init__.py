"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference algorithms on the hot path, plus the recipe that compiles the
unmodified reference op library (oracle/_ref/_ext.so).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import anything from here -- and only as the
checker or the timed CPU baseline, never as the thing measured or shipped.  pvn3d_b200/ never does.

Parity pins:
  * pn2 (PointNet++ ops): the reference has no CPU path and no golden vectors; pn2_oracle.c is
    pinned on the GPU box against oracle/_ref/_ext.so and through tests/golden/pn2_ref_*.npz.
  * meanshift / frame poses / best_fit: pinned HERE against the reference's own Python, imported
    from /root/reference (tests/golden/make_golden_cpu.py -> tests/golden/ms_*.npz, poses_*.npz).
"""
