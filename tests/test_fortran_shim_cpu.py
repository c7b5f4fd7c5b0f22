"""fortran/pic1dp_gpu_shim.F90 cannot be compiled in this image (no Fortran compiler), so its bind(c) interfaces are
checked mechanically against include/pic1dp_gpu.h: every bound name must be declared in the header with the same number
of arguments, and each argument must agree in kind -- a C scalar passed by value needs the Fortran `value` attribute, a C
pointer needs a Fortran argument WITHOUT `value` (array or intent(in/out) scalar) or a `type(c_ptr) / type(c_funptr),
value`; base types int32_t / int64_t / double / int / uint8_t map to c_int32_t / c_int64_t / c_double / c_int / c_int8_t.
One wrong `value` attribute would otherwise only show up on a user's machine.  The struct layout of pic1dp_params is
checked field by field as well."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CTYPE = {"int32_t": "c_int32_t", "int64_t": "c_int64_t", "double": "c_double", "int": "c_int", "uint8_t": "c_int8_t",
         "uint64_t": "c_int64_t", "float": "c_float"}


def c_prototypes():
    h = open(os.path.join(ROOT, "include", "pic1dp_gpu.h")).read()
    h = re.sub(r"/\*.*?\*/", " ", h, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int|void|int64_t|const char \*)\s*(pic1dp_\w+)\s*\(([^;{]*?)\)\s*;", h, flags=re.S):
        name, args = m.group(2), " ".join(m.group(3).split())
        out = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                ptr = "*" in a or "[" in a
                base = re.sub(r"\bconst\b", "", a).replace("*", " ").strip().split()
                t = base[0]
                fn = t in ("pic1dp_real64_fn", "pic1dp_gaussian_array_fn")
                out.append((t, ptr, fn))
        protos[name] = (m.group(1), out)
    return protos


def fortran_interfaces():
    s = open(os.path.join(ROOT, "fortran", "pic1dp_gpu_shim.F90")).read()
    s = re.sub(r"&\s*\n\s*", " ", s)     # join continuation lines
    out = {}
    for m in re.finditer(r"(?:integer\((\w+)\)\s+function|subroutine)\s+(\w+)\s*\(([^)]*)\)\s*bind\(c,\s*name\s*=\s*'(\w+)'\)(.*?)end (?:function|subroutine)",
                         s, flags=re.S | re.I):
        ret, args, cname, body = m.group(1), [a.strip() for a in m.group(3).split(",") if a.strip()], m.group(4), m.group(5)
        decl = {}
        for line in body.split("\n"):
            line = line.split("!")[0].strip()
            dm = re.match(r"(integer\((\w+)\)|real\((\w+)\)|type\((\w+)\))\s*((?:,\s*[\w()]+)*)\s*::\s*(.*)", line, flags=re.I)
            if not dm:
                continue
            kind = dm.group(2) or dm.group(3) or dm.group(4)
            attrs = dm.group(5).lower()
            for v in re.split(r",\s*(?![^()]*\))", dm.group(6)):
                v = v.strip()
                nm = re.match(r"(\w+)", v).group(1)
                decl[nm.lower()] = (kind.lower(), "value" in attrs, "(" in v)
        out[cname] = (ret, [decl[a.lower()] for a in args])
    return out


def test_every_bound_procedure_matches_the_header():
    protos, ifs = c_prototypes(), fortran_interfaces()
    assert len(ifs) >= 25
    for name, (ret, fargs) in ifs.items():
        assert name in protos, f"{name} is bound in the shim but not declared in include/pic1dp_gpu.h"
        cret, cargs = protos[name]
        assert len(cargs) == len(fargs), (name, len(cargs), len(fargs))
        if cret == "int":
            assert ret and ret.lower() == "c_int", name
        for k, ((ct, cptr, cfn), (fk, fval, farr)) in enumerate(zip(cargs, fargs)):
            where = f"{name} argument {k + 1} ({ct})"
            if cfn:                                   # call-back: type(c_funptr), value
                assert fk == "c_funptr" and fval, where
            elif ct == "pic1dp_gpu_t":                # opaque handle
                assert fk == "c_ptr", where
                assert fval != (name == "pic1dp_gpu_create"), where   # create receives pic1dp_gpu_t **: by reference
            elif ct == "void":                        # void *rng_ctx
                assert fk == "c_ptr" and fval, where
            elif ct == "pic1dp_params":
                assert fk == "pic1dp_params" and not fval, where
            elif cptr and fk == "c_ptr":              # nullable pointer passed as an address: type(c_ptr), value
                assert fval, where
            elif cptr:                                # array / output scalar: by reference, matching element type
                assert not fval, where + ": C pointer but Fortran `value`"
                assert fk == CTYPE[ct], where + f": {fk} vs {CTYPE[ct]}"
            else:                                     # scalar by value
                assert fval, where + ": C scalar by value needs the Fortran `value` attribute"
                assert fk == CTYPE[ct], where + f": {fk} vs {CTYPE[ct]}"


def test_params_struct_fields_match_the_header_in_order():
    h = open(os.path.join(ROOT, "include", "pic1dp_gpu.h")).read()
    body = re.search(r"typedef struct pic1dp_params \{(.*?)\} pic1dp_params;", h, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", " ", body, flags=re.S)
    cfields = [(m.group(1), m.group(2), m.group(3)) for m in re.finditer(r"(int32_t|int64_t|double)\s+(\w+)(\[[\w]+\])?;", body)]
    s = open(os.path.join(ROOT, "fortran", "pic1dp_gpu_shim.F90")).read()
    fbody = re.search(r"type, bind\(c\)(?:, public)? :: pic1dp_params(.*?)end type pic1dp_params", s, flags=re.S | re.I).group(1)
    ffields = []
    for line in fbody.split("\n"):
        m = re.match(r"\s*(integer|real)\((\w+)\)\s*::\s*(\w+)(\([\w]+\))?", line)
        if m:
            ffields.append((m.group(2).lower(), m.group(3).lower(), m.group(4)))
    assert len(cfields) == len(ffields) and len(cfields) > 20
    dims = {"PIC1DP_MAX_MODES": "64", "PIC1DP_MAX_SPECIES": "4"}
    for (ct, cn, cdim), (fk, fn, fdim) in zip(cfields, ffields):
        assert cn.lower() == fn and CTYPE[ct] == fk, (cn, fn, ct, fk)
        cd = cdim.strip("[]") if cdim else None
        fd = fdim.strip("()") if fdim else None
        assert dims.get(cd, cd) == dims.get(fd.upper() if fd else fd, fd), (cn, cdim, fdim)
