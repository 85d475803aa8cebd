"""TEST INFRASTRUCTURE ONLY — parity oracle for the ps2 stereo block matcher.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``introtocomputervision_b200``) never does.
"""
from .oracle import (  # noqa: F401
    OracleError,
    build,
    have_ref,
    have_ref_ncc,
    narrow_i8,
    ncorr,
    ncorr_fast,
    ref_ncorr,
    ref_ssd,
    refgpu,
    set_num_threads,
    num_threads,
    ssd,
    ssd_fast,
)
