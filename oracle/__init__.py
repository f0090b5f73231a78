"""CPU oracle for the CGE scoring path -- TEST INFRASTRUCTURE ONLY (see cge_oracle.c header).

Nothing under cge_jl_b200/ may import this package.
"""
from .oracle import (  # noqa: F401
    OracleTrace, build, dist, idx, js, wgcl, wgcl_directed,
)
