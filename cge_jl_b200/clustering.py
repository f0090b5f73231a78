"""Louvain wrapper -- OUT OF SCOPE for the B200 path (stays in Julia).

The reference shells out to the ``louvain_jll`` binaries (/root/reference/src/clustering.jl:14-68),
a third-party JLL whose sources and binaries are not part of the reference tree or this image.
The name is kept so that ``parseargs`` without ``-c`` fails with a clear message instead of an
AttributeError.
"""


def louvain_clust(*_args, **_kwargs):
    raise RuntimeError(
        "louvain_clust is not part of the B200 scoring path: provide communities with -c "
        "(the reference runs the external louvain_jll binaries here, clustering.jl:14-68)"
    )
