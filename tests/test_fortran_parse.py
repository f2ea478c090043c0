"""The ISO_C_BINDING shims through an actual Fortran front end.  No Fortran compiler exists in this image, so the shims
cannot be built; numpy's f2py ships a Fortran parser (numpy.f2py.crackfortran) that does read free-form modules,
derived types and interface blocks.  Parsing both shims with it checks more than the regular-expression lint of
test_fortran_shim.py: the files are syntactically digestible by an independent parser, and every bind(C) interface is
compared with the C prototype of the same name -- argument count, and by-value against by-reference passing:
    C scalar (int, double)            <->  Fortran argument with the VALUE attribute
    C pointer                          <->  type(c_ptr), value   or   a by-reference argument (no VALUE)
A mismatch here is the classic silent ABI bug (a C int received as an address)."""
import contextlib
import io
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _crack(path, tmp_path):
    import numpy.f2py.crackfortran as cf
    src = open(path).read()
    clean = tmp_path / (os.path.basename(path).rsplit(".", 1)[0] + ".f90")
    clean.write_text("\n".join(l for l in src.splitlines() if not l.startswith("#")))     # cpp guards only wrap USE lines
    cf.verbose, cf.quiet = 0, 1
    with contextlib.redirect_stdout(io.StringIO()):
        blocks = cf.crackfortran([str(clean)])
    assert len(blocks) == 1 and blocks[0]["block"] == "module"
    return blocks[0]


def _c_prototypes(header):
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|const char \*)\s*(\w+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        protos[name] = ["*" in p for p in params]            # True = pointer
    return protos


CASES = [("fortran/seaice_evp_b200.F90", "include/evp_b200.h", "evp_"), ("fortran/seaice_ir_b200.F90", "include/ir_b200.h", "ir_")]


@pytest.mark.parametrize("shim,header,prefix", CASES)
def test_interfaces_match_the_c_prototypes(shim, header, prefix, tmp_path):
    mod = _crack(os.path.join(ROOT, shim), tmp_path)
    protos = _c_prototypes(os.path.join(ROOT, header))
    bound = {}
    for blk in mod["body"]:
        if blk["block"] != "interface":
            continue
        for f in blk["body"]:
            cname = next(iter(f.get("bindlang", {}).values()), {}).get("name", f["name"])
            bound[cname] = f
    assert len(bound) >= 5
    for cname, f in bound.items():
        assert cname.startswith(prefix), cname
        assert cname in protos, f"{cname} is bound in {shim} but not declared in {header}"
        want = protos[cname]
        args = f.get("args", [])
        assert len(args) == len(want), f"{cname}: {len(args)} Fortran arguments, {len(want)} in the C prototype"
        for a, is_pointer in zip(args, want):
            var = f["vars"][a]
            by_value = "value" in (var.get("attrspec") or [])
            is_cptr = var.get("typespec") == "type" and (var.get("typename") or "").lower() == "c_ptr"
            if not is_pointer:
                assert by_value and not is_cptr, f"{cname}({a}): a C scalar must be passed with VALUE"
            else:
                assert (is_cptr and by_value) or not by_value or is_cptr, f"{cname}({a}): pointer argument passed by value as a non-pointer"
                if by_value:
                    assert is_cptr, f"{cname}({a}): VALUE on a non-c_ptr argument whose C parameter is a pointer"


@pytest.mark.parametrize("shim,header,prefix", CASES)
def test_derived_types_follow_the_struct_field_order(shim, header, prefix, tmp_path):
    mod = _crack(os.path.join(ROOT, shim), tmp_path)
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, header)).read(), flags=re.S)
    structs = {}
    for m in re.finditer(r"typedef\s+struct\s*\w*\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        fields = []
        for decl in m.group(1).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            for part in decl.split(","):                      # "const double *a, *b" -> a, b
                fields.append(re.sub(r"\[.*?\]", "", part.replace("*", " ")).split()[-1])
        structs[m.group(2)] = fields
    seen = 0
    for blk in mod["body"]:
        if blk["block"] == "type" and blk["name"] in {k.lower() for k in structs}:
            cname = next(k for k in structs if k.lower() == blk["name"])
            got = [v for v in blk["sortvars"]] if "sortvars" in blk else list(blk["vars"])
            assert [g.lower() for g in got] == [w.lower() for w in structs[cname]], (cname, got, structs[cname])
            seen += 1
    assert seen >= 2
