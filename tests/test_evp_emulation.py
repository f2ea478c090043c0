"""The EVP library without a GPU: the SHIPPED sources mpas-seaice_b200/csrc/evp_*.cu compiled for the host
(tests/emu/evp_emu.py: triple-chevron launches rewritten into emu_submit, the inline-PTX helpers into their C++ meaning;
every kernel thread a fiber, blocks one after the other) and driven through the same C ABI and the same host code
(mpas_seaice_b200.host) as the product -- BY THE BODIES OF THE `-m gpu` TESTS THEMSELVES: every test function of the
modules below that takes the ``evp_lib`` fixture is collected here with its own parametrisation and run with the
emulated library in place of the CUDA build.  Same assertions, same tolerances (bit-exact against the oracle and the
reference-executed fixtures).

What this adds to the GPU legs: it runs in the build container (where there is no device), so a change of a kernel, of
the launch sequence or of the ABI meets the oracle before any GPU time is spent; IR_EMU_ORDER=reverse runs blocks and
threads last to first (a result that changes would be a race on the device) and fills fresh device memory with 0xFF
bytes instead of zeros (IR_EMU_POISON=1); and an AddressSanitizer build reports any
access past the end of a device array (device memory is plain calloc memory of exactly the requested size).
What it does not: timing, the memory model, the peer-to-peer exchange and more than one rank (refused by the emulated
runtime), the device's own libm (exp() here is the host's).

TEST INFRASTRUCTURE: the product never loads this library (host.LIB_PATH is the CUDA build, and it fails loudly
without one)."""
import importlib
import inspect
import os
import subprocess
import sys

import pytest

from mpas_seaice_b200 import host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
import evp_emu  # noqa: E402

MODULES = ("test_gpu_parity", "test_gpu_prepost", "test_gpu_weak", "test_golden", "test_refexec_init", "test_refexec_step",
           "test_analytic_golden", "test_closed_forms", "test_gpu_fuzz", "test_fortran_shim_executed", "test_zz_late_gpu")
# left to the GPU: the sizes that take the emulation more than a few seconds each (the same code paths run on the smaller
# meshes of the same tests), and what asserts a property of the device itself
TOO_LONG = {("test_evp_subcycles_match_oracle", "hex82", 120), ("test_evp_subcycles_match_oracle", "ico5", 120),
            ("test_partial_ice_cover_masks", "hex82", "square"), ("test_partial_ice_cover_masks", "ico5", "B"),
            ("test_full_dynamics_step_on_device", "hex82", "square"), ("test_full_dynamics_step_on_device", "ico5", "B"),
            ("test_full_dynamics_step_on_device", "ico7", "A"), ("test_pinned_host_path",),
            ("test_persistent_kernel_with_partial_cover",), ("test_state_stays_resident_across_steps",),
            ("test_evp_relaxation_device", "quad40"), ("test_momentum_solve_device", "quad40")}
# EVP_EMU_ALL=1 in the environment runs those too (about five minutes more)
if os.environ.get("EVP_EMU_ALL") == "1":
    TOO_LONG = {("test_full_dynamics_step_on_device", "ico7", "A")}
# the cases of the two child runs below (reverse thread order, AddressSanitizer): one of every kind, the smallest meshes
QUICK = ("evp_subcycles_match_oracle[hex20,7]", "evp_subcycles_match_oracle[hex82,1]", "namelist_options[hex20", "namelist_options[ico3",
         "graph_and_stream_paths_agree[hex20]", "average_variational_strain[hex20]", "weak_weak_matches_oracle[evp,hex20]",
         "weak_strain_variational_divergence_matches_oracle[hex20]", "weak_full_dynamics_step", "special_boundaries_velocity",
         "set_masks", "no_ice_anywhere", "host_max_edges", "no_ocean_stress", "pwl_basis_dense", "device_pwl_precompute_bit_exact[hex20]",
         "device_wachspress_precompute_bit_exact[ico3]", "refexec_step", "refexec_init", "zz_late_gpu.device_", "state_resident[0]", "on_one_handle[1]", "pre_subcycle_",
         "device_aggregate", "random_configurations_match_oracle[1]", "random_configurations_match_oracle[7]",
         "fortran_shim_executed", "golden.")
SKIP_NAMES = {"test_linearity_of_the_stress_divergence_at_full_size", "test_split_subcycle_counts_at_full_size",
              "test_device_hibler_strength"}         # (full size; the device's exp() against libm)


def _cases():
    out = []
    for mname in MODULES:
        m = importlib.import_module(mname)
        for name, fn in inspect.getmembers(m, inspect.isfunction):
            if not name.startswith("test_") or fn.__module__ != mname or "evp_lib" not in inspect.signature(fn).parameters:
                continue
            if name in SKIP_NAMES:
                continue
            combos = [dict()]
            for mk in getattr(fn, "pytestmark", []):
                if mk.name != "parametrize":
                    continue
                names = [n.strip() for n in mk.args[0].split(",")]
                axis = []
                for v in mk.args[1]:
                    v = getattr(v, "values", v if len(names) > 1 else (v,))
                    axis.append(dict(zip(names, v)))
                combos = [dict(c, **v) for c in combos for v in axis]
            for kw in combos:
                key = (name,) + tuple(v for v in kw.values() if isinstance(v, (str, int)) and not isinstance(v, bool))
                if key in TOO_LONG or (name,) in TOO_LONG:
                    continue
                label = ",".join(os.path.basename(v)[:-4] if isinstance(v, str) and v.endswith(".npz") else str(v)
                                 for v in kw.values() if isinstance(v, (str, int, float, bool)))
                out.append(pytest.param(fn, kw, id="%s.%s[%s]" % (mname[5:], name[5:], label)))
    return out


SMALLEST = ("evp_subcycles_match_oracle[hex20,7]", "namelist_options[hex20,evp_revised", "graph_and_stream_paths_agree[hex20]",
            "average_variational_strain[hex20]", "weak_weak_matches_oracle[evp,hex20]",
            "weak_strain_variational_divergence_matches_oracle[hex20]", "special_boundaries_velocity", "no_ice_anywhere",
            "host_max_edges", "device_pwl_precompute_bit_exact[hex20]", "device_wachspress_precompute_bit_exact[ico3]",
            "refexec_step", "zz_late_gpu.device_", "state_resident[0]", "on_one_handle[1]", "pre_subcycle_cold_start_matches_oracle[ico4", "random_configurations_match_oracle[7]")
CASES = _cases()
if os.environ.get("EVP_EMU_SUBSET") == "quick":
    CASES = [c for c in CASES if any(q in c.id for q in QUICK)]
elif os.environ.get("EVP_EMU_SUBSET") == "smallest":
    CASES = [c for c in CASES if any(q in c.id for q in SMALLEST)]


@pytest.fixture
def emu(monkeypatch):
    lib = host.load_library(evp_emu.library(asan=os.environ.get("EVP_EMU_ASAN") == "1"))
    monkeypatch.setattr(host, "_lib", lib)
    return lib


def test_emulated_library_exports_the_abi(emu):
    for name in host.EXPORTS:
        assert hasattr(emu, name), name
    assert len(CASES) > {"quick": 60, "smallest": 20}.get(os.environ.get("EVP_EMU_SUBSET"), 110)


@pytest.mark.parametrize("fn,kw", CASES)
def test_gpu_test_body_on_the_emulated_kernels(emu, fn, kw, monkeypatch, capsys):
    extra = {k: v for k, v in (("monkeypatch", monkeypatch), ("capsys", capsys)) if k in inspect.signature(fn).parameters}
    fn(emu, **kw, **extra)


def _child(env_extra, timeout, subset="quick"):
    env = dict(os.environ, EVP_EMU_SUBSET=subset, **env_extra)
    return subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", "-k", "not child_run",
                           os.path.abspath(__file__)], env=env, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


@pytest.mark.skipif("EVP_EMU_SUBSET" in os.environ, reason="this IS the child run")
def test_child_run_with_blocks_and_threads_in_reverse_order():
    """IR_EMU_ORDER=reverse (tests/emu/cuda_runtime.h): blocks and threads of every launch run last to first, the fibers of
    the cooperative launch are resumed last to first.  The results must still be the oracle's, bit for bit -- a kernel
    whose threads depend on one another within a launch (a race on the device) would not survive both orders.  That
    covers the two fused kernels with their shared-memory phases, the persistent kernel with its grid barrier, the
    tile / vertex-block compaction and the pre- / post-subcycle kernels.
    The same child also runs with IR_EMU_POISON=1: every cudaMalloc'ed byte starts as 0xFF (NaN as a double, -1 as an int)
    instead of calloc's zero, so anything that counts on fresh device memory being clear shows here (cudaMalloc does not
    clear memory on the device either)."""
    r = _child(dict(IR_EMU_ORDER="reverse", IR_EMU_POISON="1"), 900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout


@pytest.mark.skipif("EVP_EMU_SUBSET" in os.environ, reason="this IS the child run")
def test_child_run_under_address_sanitizer():
    """Device-memory bounds where there is no device: the library built with -fsanitize=address, every "device" buffer a
    calloc of exactly the requested size (evp_dev_alloc -> cudaMalloc -> calloc), so an index past the end of a device
    array -- in a kernel, a layout transform or a copy -- aborts the child.  The same build carries
    -fsanitize=undefined (no recovery) with double2 / int2 aligned as on the device (16 / 8 bytes): a vector access at a
    misaligned address -- a fault on a GPU, silent on x86 --, a signed overflow in an index expression, an out-of-range
    shift or an index past a fixed-size local array aborts it too."""
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan is not installed")
    r = _child(dict(EVP_EMU_ASAN="1", LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0"), 1500,
               subset="smallest")
    assert r.returncode == 0 and "AddressSanitizer" not in r.stderr, r.stdout[-3000:] + r.stderr[-3000:]
    assert " passed" in r.stdout and "failed" not in r.stdout
