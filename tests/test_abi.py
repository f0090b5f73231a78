"""CPU: the C-ABI library loads, exports every symbol include/cge_b200.h declares, its structs
match the ctypes mirror, and it fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from cge_jl_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cge_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cge_b200_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 13
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in cge_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == syms


def test_struct_layout_matches_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include "cge_b200.h"\n#include <stdio.h>\n#include <stddef.h>\n'
                   "int main(){printf(\"%zu %zu %zu %zu %zu\\n\", sizeof(cge_b200_problem),"
                   "sizeof(cge_b200_stats), offsetof(cge_b200_problem, n_samples),"
                   "offsetof(cge_b200_stats, lo), offsetof(cge_b200_stats, ms_sweeps));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.dirname(HEADER), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert got == [C.sizeof(_lib.Problem), C.sizeof(_lib.Stats), _lib.Problem.n_samples.offset,
                   _lib.Stats.lo.offset, _lib.Stats.ms_sweeps.offset]


def test_version():
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    _lib.load().cge_b200_version(C.byref(a), C.byref(b), C.byref(c))
    assert (a.value, b.value, c.value) == (0, 1, 0)


def test_no_cpu_fallback():
    lib = _lib.load()
    if lib.cge_b200_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.cge_b200_create(0, C.byref(h))
    assert rc == _lib.ERR_CUDA and not h.value
    assert "no CPU fallback" in _lib.last_error()
    from cge_jl_b200 import divergence as dv
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dv.Scorer(0)
    # the one-shot entry point fails the same way
    p = _lib.Problem()
    out = np.zeros(7)
    n_out = C.c_int32()
    rc = lib.cge_b200_score(C.byref(p), out.ctypes.data_as(C.POINTER(C.c_double)),
                            C.byref(n_out), None)
    assert rc == _lib.ERR_CUDA


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cge_jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "libcge_oracle" not in txt, f


@pytest.mark.parametrize("n", [1, 127, 128, 129, 10000, 200000])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_plan_partitions_the_tile_sequence(n, world):
    nb = (n + 127) // 128
    prev_end = 0
    for r in range(world):
        nt, b, e = _lib.shard_plan(n, r, world)
        assert nt == nb * (nb + 1) // 2
        assert b == prev_end and e >= b
        prev_end = e
        assert abs((e - b) - nt / world) <= 1
    assert prev_end == nt


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No silent fallback when libcge_b200.so has not been built."""
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libcge_b200.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.load()
    from cge_jl_b200 import divergence as dv
    with pytest.raises(ImportError):
        dv.Scorer(0)


def test_score_multi_rejects_bad_requests():
    lib = _lib.load()
    p = _lib.Problem()
    out = np.zeros(7)
    n_out = C.c_int32()
    rc = lib.cge_b200_score_multi(C.byref(p), 1, out.ctypes.data_as(C.POINTER(C.c_double)),
                                  C.byref(n_out), None)
    assert rc == _lib.ERR_ARG and "2..8" in _lib.last_error()
    if lib.cge_b200_device_count() == 0:
        rc = lib.cge_b200_score_multi(C.byref(p), 2, out.ctypes.data_as(C.POINTER(C.c_double)),
                                      C.byref(n_out), None)
        assert rc == _lib.ERR_CUDA


def test_handle_entry_points_reject_a_null_handle():
    """The sampler and the arithmetic self-test run on the device only: without a handle (no GPU,
    no create()) they fail with an argument error instead of computing anything on the host."""
    lib = _lib.load()
    a, b = C.c_int64(), C.c_int64()
    assert lib.cge_b200_selftest_math(None, 1024, 1, C.byref(a), C.byref(b)) == _lib.ERR_ARG
    e = np.array([1, 2], dtype=np.int64)
    o = np.zeros(4, dtype=np.int64)
    pi = C.POINTER(C.c_int64)
    rc = lib.cge_b200_sample_non_edges(None, 10, 1, e.ctypes.data_as(pi), e.ctypes.data_as(pi), 1, 0,
                                       4, 1, 7, o.ctypes.data_as(pi), o.ctypes.data_as(pi), None)
    assert rc == _lib.ERR_ARG and "sample_non_edges" in _lib.last_error()


C_CLIENT = r"""
#include <stdio.h>
#include <string.h>
#include "cge_b200.h"
/* a plain-C host: what a cgo / ccall / JNI stub does, without Python in between */
int main(int argc, char **argv) {
    int a, b, c;
    cge_b200_version(&a, &b, &c);
    int64_t nt, t0, t1, rows, cols;
    if (cge_b200_shard_plan(10000, 1, 4, &nt, &t0, &t1) != CGE_B200_OK) return 2;
    if (cge_b200_table_dims(argv[1], 0, 2, &rows, &cols) != CGE_B200_OK) return 3;
    double m[6];
    if (rows != 2 || cols != 3) return 4;
    if (cge_b200_read_table(argv[1], 0, 2, rows, cols, 1, rows, m) != CGE_B200_OK) return 5; /* column-major */
    cge_b200_problem p;
    memset(&p, 0, sizeof p);
    p.struct_size = (int32_t)sizeof p;
    double out[7];
    int32_t n_out = 0;
    int rc = cge_b200_device_count() > 0 ? CGE_B200_ERR_CUDA : cge_b200_score(&p, out, &n_out, NULL);
    printf("%d.%d.%d %lld %lld %lld %g %g %g %g %g %g %d %s\n", a, b, c, (long long)nt, (long long)t0,
           (long long)t1, m[0], m[1], m[2], m[3], m[4], m[5], rc, cge_b200_last_error());
    return 0;
}
"""


def test_a_plain_c_host_can_use_the_library(tmp_path):
    """The boundary is a C ABI: a C program includes cge_b200.h, links libcge_b200.so and calls it."""
    (tmp_path / "client.c").write_text(C_CLIENT)
    (tmp_path / "t.txt").write_text("1 2 3\n4 5 6\n")
    exe = tmp_path / "client"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.dirname(HEADER),
                           str(tmp_path / "client.c"), "-o", str(exe), "-L", libdir,
                           "-l:libcge_b200.so", f"-Wl,-rpath,{libdir}"])
    out = subprocess.check_output([str(exe), str(tmp_path / "t.txt")], text=True).split(" ", 13)
    assert out[0] == "0.1.0"
    nt, t0, t1 = (int(x) for x in out[1:4])
    assert (nt, t0, t1) == _lib.shard_plan(10000, 1, 4)
    assert [float(x) for x in out[4:10]] == [1.0, 4.0, 2.0, 5.0, 3.0, 6.0]  # column-major fill
    assert int(out[10]) == _lib.ERR_CUDA                                     # no device: no fallback
