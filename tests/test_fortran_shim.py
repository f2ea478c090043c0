"""fortran/seaice_evp_b200.F90 cannot be compiled here (no Fortran compiler in the image): check
statically that its bind(C) derived types and interface blocks mirror include/evp_b200.h -- same field
names in the same order, c_int for int, c_double for double, c_ptr for pointers -- and that every bound
name is a symbol the library exports."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F90 = open(os.path.join(ROOT, "fortran", "seaice_evp_b200.F90")).read()
HDR = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "evp_b200.h")).read(), flags=re.S)


def _c_struct(name):
    body = re.search(r"typedef struct \{([^{}]*)\} " + name + ";", HDR).group(1)
    out = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        fname = re.findall(r"[A-Za-z_0-9]+", decl)[-1]
        kind = "ptr" if "*" in decl else ("double" if decl.startswith("double") else "int")
        out.append((fname, kind))
    return out


def _f_type(name):
    body = re.search(r"type, bind\(C\), public :: " + name + r"\n(.*?)end type " + name, F90, flags=re.S).group(1)
    out = []
    for line in body.strip().split("\n"):
        m = re.match(r"\s*(integer\(c_int\)|real\(c_double\)|type\(c_ptr\)) :: (\w+)", line)
        assert m, line
        out.append((m.group(2), {"integer(c_int)": "int", "real(c_double)": "double", "type(c_ptr)": "ptr"}[m.group(1)]))
    return out


def test_bind_c_types_mirror_the_header():
    for name in ("evp_mesh_desc", "evp_options", "evp_step_fields", "evp_out_fields", "evp_mesh_ext", "evp_pre_fields",
                 "evp_pre_options", "evp_post_fields", "evp_weak_mesh", "evp_weak_fields"):
        assert _f_type(name) == _c_struct(name), name


def test_bound_names_are_exported(evp_lib):
    bound = re.findall(r'bind\(C, name="(\w+)"\)', F90)
    assert len(bound) >= 12
    for n in bound:
        assert hasattr(evp_lib, n), n
    # argument counts of the interfaces agree with the prototypes
    for n in bound:
        proto = re.search(r"\b" + n + r"\s*\(([^)]*)\)", HDR).group(1)
        n_c = 0 if proto.strip() in ("", "void") else len(proto.split(","))
        iface = re.search(r"function " + n + r"\(([^)]*)\)", F90).group(1)
        n_f = 0 if not iface.strip() else len(iface.split(","))
        assert n_c == n_f, (n, n_c, n_f)


def test_enum_values_agree():
    for name in ("EVP_CR_EVP", "EVP_CR_EVP_REVISED", "EVP_CR_LINEAR", "EVP_CR_NONE", "EVP_OCEAN_QUADRATIC",
                 "EVP_OCEAN_LINEAR", "EVP_FLAG_PIN_HOST", "EVP_FLAG_OVERLAP_HALO", "EVP_SCHEME_VARIATIONAL",
                 "EVP_SCHEME_WEAK", "EVP_OK"):
        c = int(re.search(name + r"\s*=\s*(\d+)", HDR).group(1))
        f = int(re.search(name + r"\s*=\s*(\d+)", F90).group(1))
        assert c == f, name


def _code_lines():
    """Source lines without comments and preprocessor lines, continuations joined."""
    out, cur = [], ""
    for raw in F90.split("\n"):
        line = raw.split("!")[0].rstrip() if not raw.lstrip().startswith("!") else ""
        if raw.lstrip().startswith("#") or not line.strip():
            continue
        if line.rstrip().endswith("&"):
            cur += line.rstrip()[:-1] + " "
            continue
        out.append((cur + line).strip())
        cur = ""
    assert cur == ""
    return out


def test_free_form_line_length_and_block_balance():
    """No Fortran compiler here: at least the block structure must balance and no line may exceed the 132 columns of
    free-form source."""
    for i, raw in enumerate(F90.split("\n"), 1):
        assert len(raw) <= 132, f"line {i} has {len(raw)} columns"
    lines = [l.lower() for l in _code_lines()]
    opens = dict(subroutine=0, function=0, interface=0, module=0, do=0)
    ifs = types = 0
    for l in lines:
        w = l.split()
        if re.match(r"^(end\s*subroutine)\b", l): opens["subroutine"] -= 1
        elif re.match(r"^(end\s*function)\b", l): opens["function"] -= 1
        elif re.match(r"^(end\s*interface)\b", l): opens["interface"] -= 1
        elif re.match(r"^(end\s*module)\b", l): opens["module"] -= 1
        elif re.match(r"^(end\s*do)\b", l): opens["do"] -= 1
        elif re.match(r"^(end\s*if)\b", l): ifs -= 1
        elif re.match(r"^(end\s*type)\b", l): types -= 1
        elif w[0] == "subroutine": opens["subroutine"] += 1
        elif w[0] == "function" or re.match(r"^.*\bfunction\s+\w+\s*\(", l) and "bind(c" in l: opens["function"] += 1
        elif w[0] == "interface": opens["interface"] += 1
        elif w[0] == "module": opens["module"] += 1
        elif re.match(r"^do\b", l): opens["do"] += 1
        elif re.match(r"^if\s*\(.*\)\s*then$", l) or re.match(r"^else\s*if\s*\(.*\)\s*then$", l) and False: ifs += 1
        elif re.match(r"^type\s*,", l): types += 1
    assert all(v == 0 for v in opens.values()), opens
    assert ifs == 0 and types == 0, (ifs, types)


def test_every_called_c_function_is_bound_and_every_c_loc_target_is_declared():
    bound = set(re.findall(r'bind\(C, name="(\w+)"\)', F90))
    called = set(re.findall(r"\b(evp_[a-z_0-9]+)\s*\(", "\n".join(_code_lines())))
    types = {"evp_mesh_desc", "evp_options", "evp_step_fields", "evp_out_fields", "evp_mesh_ext", "evp_pre_fields",
             "evp_pre_options", "evp_post_fields", "evp_weak_mesh", "evp_weak_fields"}
    assert called - types - {"evp_b200_check"} <= bound, called - types - bound
    # inside every routine, each c_loc(x) argument is a declared pointer / target
    src = "\n".join(_code_lines())
    for m in re.finditer(r"subroutine\s+(\w+)\s*\(.*?end subroutine \1", src, flags=re.S | re.I):
        body = m.group(0)
        decl = " ".join(l for l in body.split("\n") if "::" in l)
        for var in set(re.findall(r"c_loc\((\w+)\)", body)):
            assert re.search(r"\b" + var + r"\b", decl), (m.group(1), var)


# ---------------------------------------------------------------------------------------------------------------
# fortran/seaice_ir_b200.F90 against include/ir_b200.h (incremental-remapping transport)
# ---------------------------------------------------------------------------------------------------------------

IR_F90 = open(os.path.join(ROOT, "fortran", "seaice_ir_b200.F90")).read()
IR_HDR = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "ir_b200.h")).read(), flags=re.S)


def _ir_c_struct(name):
    body = re.search(r"typedef struct " + name + r" \{([^{}]*)\} " + name + ";", IR_HDR).group(1)
    out = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        base = "double" if re.match(r"(const\s+)?double\b", decl) else "int"
        for item in re.sub(r"^(const\s+)?(int|double)\s*", "", decl).split(","):
            item = item.strip()
            m = re.match(r"(\*?)\s*(\w+)(?:\[(\d+)\])?$", item)
            assert m, item
            kind = "ptr" if m.group(1) else base
            for _ in range(int(m.group(3) or 1)):
                out.append((m.group(2), kind))
    return out


def _ir_f_type(name):
    body = re.search(r"type, bind\(C\) :: " + name + r"\n(.*?)end type " + name, IR_F90, flags=re.S).group(1)
    out = []
    for line in body.strip().split("\n"):
        m = re.match(r"\s*(integer\(c_int\)|real\(c_double\)|type\(c_ptr\)) :: (.*)$", line)
        assert m, line
        kind = {"integer(c_int)": "int", "real(c_double)": "double", "type(c_ptr)": "ptr"}[m.group(1)]
        for item in m.group(2).split(","):
            mm = re.match(r"\s*(\w+)(?:\((\d+)\))?\s*$", item)
            assert mm, item
            for _ in range(int(mm.group(2) or 1)):
                out.append((mm.group(1), kind))
    return out


def test_ir_bind_c_types_mirror_the_header():
    for name in ("ir_mesh_desc", "ir_tracer_desc", "ir_check_report", "ir_upwind_var"):
        assert _ir_f_type(name) == _ir_c_struct(name), name


def test_ir_bound_names_are_exported_with_matching_argument_counts():
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "mpas-seaice_b200", "csrc", "libir_b200.so"))
    bound = re.findall(r'bind\(C, name="(\w+)"\)', IR_F90)
    assert set(bound) == {"ir_create", "ir_set_tracers", "ir_run", "ir_set_checks", "ir_fetch_check_report",
                          "ir_fetch_conservation_sums", "ir_set_upwind_mesh", "ir_run_upwind", "ir_destroy",
                          "ir_last_error_string"}
    for n in bound:
        assert hasattr(lib, n), n
        proto = re.search(r"\b" + n + r"\s*\(([^)]*)\)", IR_HDR).group(1)
        n_c = 0 if proto.strip() in ("", "void") else len(proto.split(","))
        iface = re.search(r"function " + n + r"\(([^)]*)\)", IR_F90).group(1)
        n_f = 0 if not iface.strip() else len(iface.split(","))
        assert n_c == n_f, (n, n_c, n_f)
    assert int(re.search(r"IR_OK\s*=\s*(\d+)", IR_HDR).group(1)) == int(re.search(r"IR_OK\s*=\s*(\d+)", IR_F90).group(1))


def test_ir_shim_layout_checks():
    for i, raw in enumerate(IR_F90.split("\n"), 1):
        assert len(raw) <= 132, f"line {i} has {len(raw)} columns"
    raw_code = [l.split("!")[0].strip().lower() for l in IR_F90.split("\n") if not l.lstrip().startswith("!")]
    code, cur = [], ""
    for l in raw_code:                      # join continuation lines
        if not l:
            continue
        if l.endswith("&"):
            cur += l[:-1] + " "
            continue
        code.append((cur + l).strip())
        cur = ""
    for opener, closer in (("subroutine", "end subroutine"), ("module", "end module"), ("interface", "end interface")):
        n_open = sum(1 for l in code if re.match(r"^" + opener + r"\b", l))
        n_close = sum(1 for l in code if re.match(r"^" + closer + r"\b", l))
        assert n_open == n_close, (opener, n_open, n_close)
    assert sum(1 for l in code if re.match(r"^do\b", l)) == sum(1 for l in code if re.match(r"^end\s*do\b", l))
    n_if = sum(1 for l in code if re.match(r"^if\s*\(.*\)\s*then$", l))
    assert n_if == sum(1 for l in code if re.match(r"^end\s*if\b", l))
    # every pool array handed to c_loc is declared as a pointer in its routine
    src = "\n".join(code)
    for m in re.finditer(r"subroutine\s+(\w+)\s*\(.*?end subroutine \1", src, flags=re.S):
        body = m.group(0)
        decl = " ".join(l for l in body.split("\n") if "::" in l)
        for var in set(re.findall(r"c_loc\((\w+)\)", body)):
            assert re.search(r"\b" + var.lower() + r"\b", decl), (m.group(1), var)
