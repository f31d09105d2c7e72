"""oracle — CPU checker for the baseband->power hot path.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never from the product
package.  PARITY UNPINNED — the reference holds no implementation and no
vectors for this path (see oracle/b2p_oracle.c header).
"""
from .b2p_oracle import (  # noqa: F401
    Geometry,
    accumulate,
    accumulate_f32_naive,
    accumulate_f64,
    accumulate_omp,
    build,
    finish,
    lib,
    max_threads,
    synth_fill,
)
