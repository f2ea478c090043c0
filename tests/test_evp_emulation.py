"""The EVP library without a GPU: the SHIPPED sources mpas-seaice_b200/csrc/evp_*.cu compiled for the host
(tests/emu/evp_emu.py: triple-chevron launches rewritten into emu_submit, the inline-PTX helpers into their C++ meaning;
every kernel thread a fiber, blocks one after the other) and driven through the same C ABI and the same host code
(mpas_seaice_b200.host) as the product -- by calling the bodies of the `-m gpu` tests themselves.

What this adds to the GPU legs: it runs in the build container (where there is no device), so a change of a kernel, of
the launch sequence or of the ABI is checked against the oracle before any GPU time is spent; IR_EMU_ORDER=reverse runs
blocks and threads last to first (a result that changes would be a race on the device); and an AddressSanitizer build
reports any access past the end of a device array (device memory is plain calloc memory of exactly the requested size).
What it does not: timing, the memory model, the persistent cooperative kernel and the peer-to-peer exchange (both
refused by the emulated runtime), more than one rank.

TEST INFRASTRUCTURE: the product never loads this library (host.LIB_PATH is the CUDA build, and it fails loudly
without one)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from mpas_seaice_b200 import host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
import evp_emu  # noqa: E402


@pytest.fixture
def emu(monkeypatch):
    lib = host.load_library(evp_emu.library())
    monkeypatch.setattr(host, "_lib", lib)
    monkeypatch.setenv("EVP_B200_PERSISTENT", "0")
    return lib


def test_emulated_library_exports_the_abi(emu):
    for name in host.EXPORTS:
        assert hasattr(emu, name), name


@pytest.mark.parametrize("kind,nsub", [("hex20", 1), ("hex20", 7), ("ico3", 5)])
def test_subcycles_match_oracle(emu, kind, nsub):
    import test_gpu_parity as t
    t.test_evp_subcycles_match_oracle(emu, kind, nsub)


def test_late_reference_executed_fixtures_on_the_emulated_kernels(emu):
    """tests/test_zz_late_gpu.py (ice shelves, the 'fekete' / dunavant-12 rules): the device legs that have not met a
    B200 yet, here against the emulated kernels -- reference-executed outputs, bit for bit."""
    import test_zz_late_gpu as t
    import test_refexec_init
    assert len(t.FILES) == 2 and len(test_refexec_init.CPU_FILES) == 3
    for path in t.FILES:
        t.test_device_reproduces_the_reference_executed_step_with_ice_shelves(emu, path)
    for path in test_refexec_init.CPU_FILES:
        t.test_device_precompute_reproduces_the_reference_executed_arrays_of_the_late_rules(emu, path)
