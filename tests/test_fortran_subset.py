"""The Fortran-subset interpreter (tests/golden/fortran_subset.py) that produces the reference-executed golden vectors:
its arithmetic must be the source's, one IEEE operation per operator, in Fortran's precedence and argument-association
rules.  These are the rules the fixtures' claim rests on; /root/reference is not needed here."""
import math
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import fortran_subset as F  # noqa: E402


def _interp(src, tmp_path, defined=()):
    p = tmp_path / "m.F"
    p.write_text(src)
    I = F.Interpreter(defined=defined)
    I.load(str(p))
    return I


def _eval(expr, **env):
    I = F.Interpreter()
    fr = F.Frame(I, F.Sub("t", [], [], ""))
    fr.vars.update(env)
    return I.ev(F.parse_expr(expr), fr)


def test_expression_semantics():
    a, b, c = 0.1, 0.7, 3.3
    assert _eval("a + b * c", a=a, b=b, c=c) == a + (b * c)
    assert _eval("a * b * c", a=a, b=b, c=c) == (a * b) * c                    # left to right
    assert _eval("a / b * c", a=a, b=b, c=c) == (a / b) * c
    assert _eval("a - b - c", a=a, b=b, c=c) == (a - b) - c
    assert _eval("-a**2", a=a) == -(a * a)                                      # ** binds tighter than unary minus
    assert _eval("a**2", a=a) == a * a and _eval("a**3", a=a) == a * (a * a)    # integer powers are multiplications
    assert _eval("2**3**2") == 512                                              # ** is right-associative
    assert _eval("a**0.5", a=a) == math.pow(a, 0.5)
    assert _eval("7 / 2") == 3 and _eval("-7 / 2") == -3 and _eval("7.0 / 2") == 3.5   # integer division truncates
    assert _eval("1.0e-11_RKIND") == 1.0e-11 and _eval("0.5_RKIND * 4") == 2.0 and _eval("1.5d0") == 1.5
    assert _eval("sign(1.0_RKIND, x)", x=-0.0) == -1.0 and _eval("sign(1.0_RKIND, x)", x=2.0) == 1.0
    assert _eval("max(a, b, c)", a=a, b=b, c=c) == c and _eval("min(a,b)", a=a, b=b) == a
    assert _eval("sqrt(a*a + (b*b + c*c) / 4.0_RKIND)", a=a, b=b, c=c) == math.sqrt(a * a + (b * b + c * c) / 4.0)
    assert _eval("a > b .or. .not. (a >= b) .and. c == c", a=a, b=b, c=c) is True
    assert _eval("a .lt. b", a=a, b=b) is True and _eval("a /= b", a=a, b=b) is True
    assert _eval("nint(x)", x=2.5) == 3 and _eval("nint(x)", x=-2.5) == -3 and _eval("nint(x)", x=0.49) == 0   # half away from zero
    assert _eval("int(x)", x=-2.7) == -2 and _eval("mod(-7, 3)") == -1 and _eval("modulo(-7, 3)") == 2      # sign of the dividend / divisor
    assert _eval("mod(x, 2.0_RKIND)", x=-5.5) == -1.5 and _eval("real(7, RKIND) / 2") == 3.5
    assert _eval("abs(x)", x=-0.0) == 0.0 and math.copysign(1.0, _eval("abs(x)", x=-0.0)) == 1.0
    assert _eval("merge(a, b, a > b)", a=a, b=b) == b
    # a sum written on one line is added left to right: not the same bits as another association
    x = [1.0e16, 1.0, -1.0e16, 1.0]
    assert _eval("p + q + r + s", p=x[0], q=x[1], r=x[2], s=x[3]) == ((x[0] + x[1]) + x[2]) + x[3] == 1.0
    assert _eval("p + (q + r) + s", p=x[0], q=x[1], r=x[2], s=x[3]) == 1.0
    assert _eval("p + r + (q + s)", p=x[0], q=x[1], r=x[2], s=x[3]) == 2.0


def test_arrays_are_one_based_with_the_dimensions_reversed():
    a = np.arange(24, dtype=np.float64).reshape(4, 3, 2)       # Fortran a(2,3,4)
    fa = F.FArray(a)
    assert fa.get((1, 1, 1)) == 0.0 and fa.get((2, 3, 4)) == 23.0 and fa.get((2, 1, 1)) == 1.0 and fa.get((1, 2, 1)) == 2.0
    fa.set((slice(None, None), 2, 3), 7.0)
    assert np.all(a[2, 1, :] == 7.0) and a[2, 0, 0] == 12.0
    with pytest.raises(F.FortranError):
        fa.get((3, 1, 1))
    assert F._sum(F.FArray(np.array([[1.0e16, 1.0], [-1.0e16, 1.0]]).T.copy())) == 1.0     # element order


SRC = """
module m
  integer, parameter :: TWO = 2
  real(kind=RKIND), parameter :: half = 1.0_RKIND / TWO, quarter = half**2
contains
  subroutine kernel(x, y, n, total)
    real(kind=RKIND), dimension(:), intent(inout) :: x
    real(kind=RKIND), dimension(:), intent(in) :: y
    integer, intent(in) :: n
    real(kind=RKIND), intent(out) :: total
    real(kind=RKIND), dimension(2,2) :: w
    integer :: i
    total = 0.0_RKIND
    w(1,2) = 3.0_RKIND
#ifdef FAST
    total = -1.0_RKIND
#else
!$omp parallel do
    do i = 1, n
       if (y(i) < 0.0_RKIND) cycle
       call scale(x(i), &
            y(i))            ! array element by reference
       total = total + x(i) * w(1,2)
       if (i == 4) exit
    enddo
#endif
  end subroutine kernel

  subroutine scale(a, b)
    real(kind=RKIND), intent(inout) :: a
    real(kind=RKIND), intent(in) :: b
    a = a * b + quarter
  end subroutine scale

  subroutine driver(domain)
    type(domain_type) :: domain
    type(block_type), pointer :: block
    real(kind=RKIND), dimension(:), pointer :: x, y
    integer, pointer :: n
    logical, pointer :: on
    real(kind=RKIND) :: t
    block => domain % blocklist
    do while (associated(block))
       call MPAS_pool_get_config(block % configs, "config_on", on)
       call MPAS_pool_get_dimension(block % dimensions, "n", n)
       call MPAS_pool_get_subpool(block % structs, "state", statePool)
       call MPAS_pool_get_array(statePool, "x", x)
       call MPAS_pool_get_array(statePool, "y", y)
       if (on) then
          call kernel(x, y, n, t)
          result = t
       else if (.not. on) then
          result = -5.0_RKIND
       endif
       call mpas_timer_stop("k")
       block => block % next
    end do
  end subroutine driver
end module m
"""


def test_statements_calls_and_the_pool_mapping(tmp_path):
    import types
    I = _interp(SRC, tmp_path)
    assert I.globals["two"] == 2 and I.globals["half"] == 0.5 and I.globals["quarter"] == 0.25
    x = np.array([1.0, 2.0, 3.0, 4.0, 5.0, 0.0])
    y = np.array([0.5, -1.0, 2.0, 0.1, 9.0, 0.0])
    I.pool.update({"x": F.FArray(x), "y": F.FArray(y), "n": 5, "config_on": True})
    I.globals["result"] = 0.0
    I.noop.add("mpas_timer_stop")
    block = types.SimpleNamespace(structs=1, configs=2, dimensions=3, next=None)
    I.call("driver", types.SimpleNamespace(blocklist=block))
    want_x = [1.0 * 0.5 + 0.25, 2.0, 3.0 * 2.0 + 0.25, 4.0 * 0.1 + 0.25, 5.0]          # i = 2 cycled, exit after i = 4
    assert list(x[:5]) == want_x
    assert I.globals["result"] == ((0.0 + want_x[0] * 3.0) + want_x[2] * 3.0) + want_x[3] * 3.0
    assert I.trace == ["driver", "kernel", "scale", "scale", "scale"]
    I.pool["config_on"] = False
    I.call("driver", types.SimpleNamespace(blocklist=block))
    assert I.globals["result"] == -5.0
    # an unknown call is an error, never skipped
    I.noop.clear()
    with pytest.raises(F.FortranError):
        I.call("driver", types.SimpleNamespace(blocklist=block))
    # the preprocessor picks the other branch when the macro is defined
    J = _interp(SRC, tmp_path, defined=("FAST",))
    fr = J.call("kernel", F.FArray(x), F.FArray(y), 5, 0.0)
    assert fr.vars["total"] == -1.0


def test_reference_executed_fixtures_say_where_they_come_from():
    """The refexec_*.npz fixtures replayed by tests/test_golden.py carry their provenance: the list of reference
    subroutines the interpreter executed to produce the outputs."""
    import glob
    files = sorted(glob.glob(os.path.join(HERE, "golden", "refexec_*.npz")))
    assert len(files) >= 6
    seen = set()
    for f in files:
        prov = str(np.load(f)["provenance"])
        assert "interpreting the reference's Fortran source" in prov
        seen |= {w.strip() for w in prov.split(":", 1)[1].split(",")}
    for name in ("subcycle_velocity_solver", "single_subcycle_velocity_solver", "seaice_internal_stress",
                 "seaice_strain_tensor_variational", "seaice_stress_tensor_variational", "seaice_stress_divergence_variational",
                 "seaice_evp_constitutive_relation", "seaice_evp_constitutive_relation_revised",
                 "seaice_linear_constitutive_relation", "seaice_average_strains_on_vertex", "ocean_stress_coefficient",
                 "solve_velocity", "solve_velocity_revised"):
        assert name in seen, name


def test_string_literals_keep_their_case_in_conditions(tmp_path):
    src = """
module m
contains
  subroutine pick(name, out)
    character(len=*), intent(in) :: name
    integer, intent(out) :: out
    out = 0
    if (trim(name) == 'iceAreaCategory') then
       out = 1
    elseif (trim(name) == 'iceVolumeCategory' .or. &
            trim(name) == 'snowVolumeCategory') then
       out = 2
    endif
    if (trim(name) == 'surfaceTemperature') out = 3
  end subroutine pick
end module m
"""
    I = _interp(src, tmp_path)
    for name, want in (("iceAreaCategory", 1), ("snowVolumeCategory", 2), ("surfaceTemperature", 3), ("iceareacategory", 0)):
        assert I.call("pick", name + "   ", 0).vars["out"] == want
