"""bind(C) bridge for tests/golden/fortran_subset.py: lets INTERPRETED Fortran call a real C library.

TEST INFRASTRUCTURE.  The ISO_C_BINDING shims under fortran/ cannot be compiled in this image (no Fortran compiler), so
their interface blocks and bind(C) types could only be linted.  With this bridge the interpreter EXECUTES them: every
`function f(...) bind(C, name="f")` found in an interface block of a loaded file becomes a call into the shared library
through ctypes, marshalled from the Fortran declarations themselves -- VALUE scalars by value, everything else by
reference, `type(c_ptr)` as an address (c_loc of an array = the address of its first element), bind(C) derived types as
C structs with the components in declaration order, `dimension(*)` arguments as arrays.  A mismatch between a shim's
interface and the library's C prototype therefore fails the way it would at run time in a compiled host: wrong results or
a crash, not a lint message.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from fortran_subset import FArray, FortranError


class CBridge:
    def __init__(self, interp, lib):
        self.I, self.lib = interp, lib
        self.keep = []
        self._structs = {}
        self.calls = []                       # names of the C functions called, in order
        self.returns = []                     # (name, return value)
        for name in interp.cfuncs:
            if hasattr(lib, name):
                interp.func_hooks[name] = self._make(name)

    # ---- derived types <-> C structs
    def struct_class(self, typename):
        if typename not in self._structs:
            fields = []
            for comp, kind, dims in self.I.types[typename]:
                ct = {"int": C.c_int, "real": C.c_double, "ptr": C.c_void_p, "log": C.c_int}[kind]
                if dims is not None:
                    ct = ct * int(dims)
                fields.append((comp, ct))
            self._structs[typename] = type("c_" + typename, (C.Structure,), {"_fields_": fields})
        return self._structs[typename]

    def address(self, v):
        if v is None:
            return None
        if isinstance(v, (int, np.integer)):
            return int(v) or None
        if isinstance(v, FArray):
            a = v.a
            if not a.flags["C_CONTIGUOUS"]:
                raise FortranError("c_loc of a non-contiguous array section")
            self.keep.append(a)
            return a.ctypes.data
        raise FortranError("cannot take the address of %r" % (v,))

    def to_struct(self, obj):
        cls = self.struct_class(obj._type)
        st = cls()
        for comp, kind, dims in self.I.types[obj._type]:
            val = getattr(obj, comp)
            if dims is not None:
                arr = getattr(st, comp)
                for i in range(int(dims)):
                    x = val.a.reshape(-1)[i]
                    arr[i] = (self.address(x) or 0) if kind == "ptr" else x
            elif kind == "ptr":
                setattr(st, comp, self.address(val))
            elif kind == "log":
                setattr(st, comp, int(bool(val)))
            else:
                setattr(st, comp, val)
        return st

    def from_struct(self, st, obj):
        for comp, kind, dims in self.I.types[obj._type]:
            if dims is None and kind in ("int", "real"):
                setattr(obj, comp, getattr(st, comp))

    # ---- calls
    def _make(self, name):
        def hook(interp, fr, args):
            cf = interp.cfuncs[name]
            if len(args) != len(cf["args"]):
                raise FortranError("%s called with %d arguments, its interface has %d" % (name, len(args), len(cf["args"])))
            cargs, after = [], []
            for dummy, (kw, node) in zip(cf["args"], args):
                d = cf["decl"][dummy]
                kind = d["kind"]
                if kind == "ptr":
                    if d["value"]:
                        cargs.append(C.c_void_p(self.address(interp.ev(node, fr))))
                    else:                                   # type(c_ptr), intent(out): the address comes back
                        holder = C.c_void_p()
                        cargs.append(C.byref(holder))
                        after.append((node, holder))
                elif kind in ("int", "real"):
                    ct = C.c_int if kind == "int" else C.c_double
                    if d["value"]:
                        cargs.append(ct(interp.ev(node, fr)))
                    elif d["array"]:
                        v = interp.ev(node, fr)
                        want = np.int32 if kind == "int" else np.float64
                        if v.a.dtype != want:
                            raise FortranError("%s: argument %s is %s, the interface says %s" % (name, dummy, v.a.dtype, want))
                        cargs.append(C.c_void_p(self.address(v)))
                    else:
                        holder = ct(interp.ev(node, fr) if not d["out"] else 0)
                        cargs.append(C.byref(holder))
                        if d["out"]:
                            after.append((node, holder))
                elif kind.startswith("type:"):
                    v = interp.ev(node, fr)
                    if d["array"]:
                        objs = list(v.a.reshape(-1))
                        arr = (self.struct_class(kind[5:]) * len(objs))(*[self.to_struct(o) for o in objs])
                        self.keep.append(arr)
                        cargs.append(arr)
                    else:
                        st = self.to_struct(v)
                        self.keep.append(st)
                        cargs.append(C.byref(st))
                        if d["out"]:
                            after.append((v, st))
                else:
                    raise FortranError("%s: cannot marshal argument %s of kind %s" % (name, dummy, kind))
            fn = getattr(self.lib, name)
            fn.restype = C.c_void_p if cf["result"] == "ptr" else C.c_int
            self.calls.append(name)
            rc = fn(*cargs)
            self.returns.append((name, rc))
            for target, holder in after:
                if isinstance(holder, C.Structure):
                    self.from_struct(holder, target)
                else:
                    interp._assign(target, holder.value, fr)
            return rc
        return hook
