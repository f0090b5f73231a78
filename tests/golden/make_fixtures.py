"""Regenerates tests/golden/*.npz from the reference's example and test data.

Run in the BUILD container only (it reads /root/reference, which does not exist on the GPU
box):  python tests/golden/make_fixtures.py

Each .npz holds exactly what the reference's ``parseargs`` returns for the corresponding
command line (1-based ``edges``/``comm``), produced by this repo's mirror of ``parseargs``
(cge_jl_b200/auxilary.py).  The three embedding formats of the 115-node test graph parse to
the same matrix, which the script checks.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cge_jl_b200.auxilary import parseargs  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def save(name, tup):
    edges, ew, vw, comm, _clusters, emb = tup[:6]
    np.savez_compressed(os.path.join(OUT, name), edges=edges, eweights=ew, vweights=vw,
                        comm=comm, embedding=emb)
    print(name, edges.shape, comm.shape, emb.shape)


def main():
    t = f"{REF}/test"
    a = parseargs(["-g", f"{t}/test.edgelist", "-c", f"{t}/test1col.ecg", "-e",
                   f"{t}/test_n2v.embedding"])
    b = parseargs(["-g", f"{t}/test.edgelist", "-c", f"{t}/test2col.ecg", "-e",
                   f"{t}/test_ordered.embedding"])
    c = parseargs(["-g", f"{t}/test_weights.edgelist", "-c", f"{t}/test2col.ecg", "-e",
                   f"{t}/test_unordered.embedding"])
    for x in (b, c):
        assert np.array_equal(a[0], x[0]) and np.array_equal(a[3], x[3])
        assert np.array_equal(a[5], x[5])
    assert np.allclose(c[1], 1.42)
    save("test115.npz", a)
    save("test115_weighted.npz", c)
    e = f"{REF}/example"
    save("example10k.npz", parseargs(["-g", f"{e}/10k.edgelist", "-c", f"{e}/10k.ecg", "-e",
                                      f"{e}/10k.embedding", "--force-exact"]))


if __name__ == "__main__":
    main()
