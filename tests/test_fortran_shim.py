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
