"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/pic1dp_gpu.h
declares, the ctypes mirror of the structs matches the header, and the product fails loudly without a GPU
(no compute calls are made here)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pic1dp_gpu.h")


def _declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pic1dp_(?:gpu|host)_\w+)\s*\(", txt)))


def test_header_symbols_all_exported(capi):
    from pic1dp_b200 import _capi
    declared = _declared_symbols()
    assert len(declared) >= 20
    assert sorted(_capi.EXPORTS) == declared, "ctypes EXPORTS list and the header disagree"
    for name in declared:
        assert hasattr(capi, name), f"libpic1dp_b200.so does not export {name}"


def test_exported_symbols_are_plain_c(capi):
    from pic1dp_b200 import _capi
    out = subprocess.run(["nm", "-D", "--defined-only", _capi.lib_path()], capture_output=True, text=True).stdout
    names = {l.split()[-1] for l in out.splitlines() if " T " in l}
    for name in _declared_symbols():
        assert name in names  # unmangled => extern "C"


def test_struct_layout_matches_header(capi, tmp_path):
    """Compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirror."""
    from pic1dp_b200 import _capi
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "pic1dp_gpu.h"\nint main(void){\n'
                   'printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(pic1dp_params), offsetof(pic1dp_params, lx),'
                   'offsetof(pic1dp_params, charge), offsetof(pic1dp_params, iptcldist), offsetof(pic1dp_params, capacity),'
                   'offsetof(pic1dp_params, fuse), sizeof(pic1dp_counters));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(t) for t in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    P = _capi.Params
    want = [C.sizeof(P), P.lx.offset, P.charge.offset, P.iptcldist.offset, P.capacity.offset, P.fuse.offset,
            C.sizeof(_capi.Counters)]
    assert got == want


def test_defaults_are_the_reference_input_file(capi):
    import pic1dp_b200 as P
    p = P.default_params()
    # /root/reference/src/pic1dp_input.F90:43-138
    assert (p.nx, p.nmode, p.modes[0], p.nspecies) == (192, 1, 1, 1)
    assert p.lx == 2.0 * 3.1415926535897932384626 / 0.36 and p.dt == 0.05
    assert (p.charge[0], p.mass[0], p.temperature[0], p.temperature2[0], p.density[0], p.v0[0]) == (-1.0, 1.0, 1.0, 1.0, 0.9, 5.0)
    assert (p.iptcldist, p.deltaf, p.linear, p.iptclshape) == (3, 1, 0, 4)
    assert p.capacity == 6400000 and capi.pic1dp_gpu_abi_version() == p.abi_version == 2
    assert p.struct_bytes == C.sizeof(P.Params)


def test_error_strings(capi):
    assert capi.pic1dp_gpu_strerror(0) == b"ok"
    assert b"CPU fallback" in capi.pic1dp_gpu_strerror(7)
    assert capi.pic1dp_gpu_strerror(99) == b"unknown error"


@pytest.mark.parametrize("bad", [dict(nx=1), dict(nmode=0), dict(nmode=65), dict(nspecies=5), dict(lx=-1.0),
                                 dict(iptcldist=4), dict(iptclshape=0), dict(linear=1, deltaf=0), dict(capacity=0),
                                 dict(rank=2, nranks=2), dict(mass=[0.0]), dict(modes=[0]), dict(deposit_mode=7),
                                 dict(abi_version=1), dict(arith_mode=2)])
def test_create_rejects_invalid_parameters_before_touching_the_gpu(bad):
    """input_init-style validation (src/pic1dp_input.F90:287-308) -> PIC1DP_EINVAL, no exception across the ABI."""
    import pic1dp_b200 as P
    with pytest.raises(P.Pic1dpError) as e:
        P.Pic1dGpu(P.default_params(**bad))
    assert e.value.code == 1


def test_no_cpu_fallback():
    """On a box without a GPU the product must fail loudly, not compute on the CPU."""
    import torch
    import pic1dp_b200 as P
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(P.Pic1dpError) as e:
        P.Pic1dGpu(P.default_params(capacity=16))
    assert e.value.code == 7  # PIC1DP_ENODEVICE


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under pic1dp_b200/ may import, link or execute it."""
    pkg = os.path.join(ROOT, "pic1dp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "host.py" and "oracle" not in txt, (dirpath, f)
    out = subprocess.run(["ldd", os.path.join(pkg, "libpic1dp_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_petsc_decide_split():
    import pic1dp_b200 as P
    from oracle import oracle as O
    for n, npe in ((6400000, 4), (10, 3), (7, 8), (0, 2), (100000001, 8)):
        spans = [P.petsc_decide(n, npe, r) for r in range(npe)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for r in range(npe):
            assert spans[r] == O.petsc_decide(n, npe, r)
            if r:
                assert spans[r][0] == spans[r - 1][1]
            assert spans[r][1] - spans[r][0] == n // npe + (1 if r < n % npe else 0)


def test_cpp_host_driver_builds_and_fails_loudly_without_gpu():
    """host/pic1dp_host (C++ mirror of `program pic1dp`) links against the C ABI; with no GPU it must exit non-zero
    with the library's no-device error instead of computing on the CPU."""
    import torch
    from pic1dp_b200 import build
    exe = build.build_host()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    out = subprocess.run([exe, "nparticle_max=1000", "time_max=1"], capture_output=True, text=True)
    assert out.returncode == 1
    assert "no CUDA device" in out.stderr
    bad = subprocess.run([exe, "linear=1", "deltaf=0"], capture_output=True, text=True)  # input_init check
    assert bad.returncode == 1 and "not implemented" in bad.stderr


def test_cpp_host_multirand_against_known_answers_and_oracle():
    """The product-side RNG of host/pic1dp_host.cpp (no GPU needed): engine heads after the default seeds equal the
    reference's known-answer values (tests/golden/multirand_kat.json, from src/multirand.F90:396-425); uniform and
    Gaussian draws after the constant-seed initialisation (rank 2, warm-up 5) equal the KAT-pinned oracle bit for bit,
    including the one-value buffer of the polar method across calls."""
    import json
    from pic1dp_b200 import build
    from oracle import oracle as O
    exe = build.build_host()
    out = subprocess.run([exe, "multirand_selftest"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "multirand_kat.json")))
    heads = {1: kat["kiss64"], 2: kat["mt19937_64_head"], 3: kat["superkiss64_head"]}
    lines = {(l.split()[0], int(l.split()[1])): l.split()[2:] for l in out.stdout.splitlines() if l.strip()}
    for al in (1, 2, 3):
        assert [int(t) for t in lines[("default", al)]] == heads[al], al
        r = O.MultiRand()
        r.init_const(al, 2, 5)
        assert [float.fromhex(t) for t in lines[("real64", al)]] == [r.real64() for _ in range(5)], al
        g = list(r.gaussian_array(7)) + list(r.gaussian_array(4))
        assert [float.fromhex(t) for t in lines[("gauss", al)]] == g, al
