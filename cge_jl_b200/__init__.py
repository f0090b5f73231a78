"""cge_jl_b200 -- B200-native scoring path of CGE.jl behind the reference's entry points.

Public names are the ones ``CGE.jl`` exports (/root/reference/src/CGE.jl:11-21).  ``wGCL`` and
``wGCL_directed`` run entirely in ``libcge_b200.so`` (hand-written sm_100a CUDA behind a C ABI,
see include/cge_b200.h); there is no CPU fallback.
"""
from .auxilary import parseargs  # noqa: F401
from .clustering import louvain_clust  # noqa: F401
from .landmarks import landmarks  # noqa: F401


def __getattr__(name):
    # the scorer needs the native library; import it lazily so that parsing / landmark
    # selection stay usable for tooling that never scores
    if name in ("wGCL", "wGCL_directed", "Scorer", "draw_samples"):
        from . import divergence

        return getattr(divergence, name)
    raise AttributeError(name)
