"""CPU oracle for the CGE scoring path -- TEST INFRASTRUCTURE ONLY (see cge_oracle.c header).

Nothing under cge_jl_b200/ may import this package.
"""
from .oracle import (  # noqa: F401
    OracleMtTrace, OracleStreamTrace, OracleTrace, build, dist, host_threads, idx, js, wgcl,
    wgcl_directed, wgcl_mt, wgcl_stream,
)
