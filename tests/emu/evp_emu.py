"""Builds tests/_build/libevp_b200_emu.so: the SHIPPED sources of the EVP library (mpas-seaice_b200/csrc/evp_*.cu)
compiled for the host, every kernel thread a fiber (tests/emu/cuda_runtime.h, tests/emu/evp/cuda_runtime.h).

TEST INFRASTRUCTURE ONLY -- nothing of this is loaded by the product.  The sources are not edited for it; this script
makes a rewritten COPY under tests/_build/evp_emu/ with exactly two kinds of change:

  * ``kernel<<<grid, block, smem, stream>>>(args)``  ->  ``emu_submit(emu_cfg(grid, block, smem, stream), kernel, args)``
  * the bodies of the few helpers written in inline PTX (mbarrier, cp.async.bulk, acquire / release accesses, the global
    timer) -> their plain C++ meaning (a bulk copy is a memcpy that has completed when the call returns; a barrier wait
    therefore never waits); a helper with inline PTX that is not in the table below stops the build;
  * ``extern __shared__ ... evp_smem_raw[]``  ->  a pointer to the emulated dynamic shared memory;
  * ``cudaLaunchCooperativeKernel((const void *)kern, ...)``  ->  ``emu_launch_cooperative(kern, ...)`` (the typed pointer):
    all blocks alive at once, so the persistent kernel's grid barrier works (tests/emu/cuda_runtime.h, emu_launch_grid).

Used by tests/test_evp_emulation.py."""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "mpas-seaice_b200", "csrc")
BUILD = os.path.join(ROOT, "tests", "_build", "evp_emu")
SOURCES = ("evp_abi.cu", "evp_kernels.cu", "evp_halo.cu", "evp_precompute.cu", "evp_prepost.cu", "evp_weak.cu")
HEADERS = ("evp_internal.cuh", "evp_quadrature_tables.inc")

PTX_HELPERS = {
    "smem_u32": "{ return (uint32_t)(uintptr_t)p; }",
    "mbar_init": "{ *bar = 0; (void)count; }",
    "mbar_fence_init": "{ }",
    "mbar_expect_tx": "{ (void)bar; (void)bytes; }",
    "bulk_g2s": "{ memcpy(dst, src, bytes); (void)bar; }",
    "mbar_wait": "{ (void)bar; (void)parity; }",
    "evp_ld_acquire_gpu_u32": "{ emu_yield(); return *(const volatile unsigned *)p; }",   # polled in the grid barrier
    "evp_st_release_gpu_u32": "{ *(volatile unsigned *)p = v; }",
    "evp_ld_acquire_sys": "{ return *(const volatile int *)p; }",
    "evp_st_relaxed_sys": "{ *(volatile int *)p = v; }",
    "evp_globaltimer": "{ return 0ull; }",
}


def _match(text, i, open_ch, close_ch):
    """index of the bracket closing the one at text[i]"""
    depth = 0
    for k in range(i, len(text)):
        if text[k] == open_ch:
            depth += 1
        elif text[k] == close_ch:
            depth -= 1
            if depth == 0:
                return k
    raise ValueError("unbalanced %s at %d" % (open_ch, i))


def rewrite_ptx_helpers(text, name):
    for fn, body in PTX_HELPERS.items():
        for m in list(re.finditer(r"\b%s\s*\(" % re.escape(fn), text))[::-1]:
            close = _match(text, m.end() - 1, "(", ")")
            k = close + 1
            while k < len(text) and (text[k] in " \t\n" or text.startswith("//", k)):
                k = text.index("\n", k) + 1 if text.startswith("//", k) else k + 1
            if k < len(text) and text[k] == "{":                     # a definition, not a call
                end = _match(text, k, "{", "}")
                if "asm" in text[k:end]:
                    text = text[:k] + body + text[end + 1:]
    left = re.search(r"\basm\b", text)
    if left:
        line = text.count("\n", 0, left.start()) + 1
        raise RuntimeError("%s:%d: inline PTX outside the helpers tests/emu/evp_emu.py knows" % (name, line))
    return text


def rewrite_launches(text, name):
    out, pos = [], 0
    while True:
        i = text.find("<<<", pos)
        if i < 0:
            out.append(text[pos:])
            break
        # the kernel expression: identifier, optionally followed by <template arguments>
        k = i
        while text[k - 1] in " \t":
            k -= 1
        if text[k - 1] == ">":
            depth, k2 = 0, k - 1
            while True:
                if text[k2] == ">":
                    depth += 1
                elif text[k2] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                k2 -= 1
            k = k2
        s = k
        while text[s - 1].isalnum() or text[s - 1] in "_:":
            s -= 1
        kernel = text[s:i].strip()
        j = text.index(">>>", i)
        cfg = text[i + 3:j]
        p = j + 3
        while text[p] in " \t\n":
            p += 1
        if text[p] != "(":
            raise RuntimeError("%s: launch of %s without an argument list" % (name, kernel))
        q = _match(text, p, "(", ")")
        args = text[p + 1:q].strip()
        out.append(text[pos:s])
        out.append("emu_submit(emu_cfg(%s), %s%s)" % (cfg, kernel, (", " + args) if args else ""))
        pos = q + 1
    return "".join(out)


def rewrite(text, name):
    text = rewrite_ptx_helpers(text, name)
    text = rewrite_launches(text, name)
    text = re.sub(r"extern\s+__shared__\s+__align__\(\d+\)\s+unsigned\s+char\s+(\w+)\[\];",
                  r"unsigned char *\1 = emu_evp_dyn_smem;", text)
    text = text.replace("__cvta_generic_to_shared", "(uintptr_t)")
    text = text.replace("cudaLaunchCooperativeKernel((const void *)kern,", "emu_launch_cooperative(kern,")
    return text


def library(asan=False):
    out = os.path.join(ROOT, "tests", "_build", "libevp_b200_emu%s.so" % ("_asan" if asan else ""))
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(HERE, "cuda_runtime.h"),
            os.path.join(HERE, "evp", "cuda_runtime.h"), os.path.join(HERE, "evp", "nccl.h"), os.path.abspath(__file__),
            os.path.join(ROOT, "include", "evp_b200.h")]
    if os.path.exists(out) and all(os.path.getmtime(p) <= os.path.getmtime(out) for p in deps):
        return out
    os.makedirs(os.path.join(BUILD, "csrc"), exist_ok=True)
    # same relative position of the public header as in the tree: the sources include "../../include/evp_b200.h"
    inc = os.path.join(BUILD, "..", "include_link")
    cpps = []
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f)) as fh:
            text = rewrite(fh.read(), f)
        text = text.replace('"../../include/evp_b200.h"', '"%s"' % os.path.join(ROOT, "include", "evp_b200.h"))
        dst = os.path.join(BUILD, "csrc", f[:-3] + ".cpp" if f.endswith(".cu") else f)
        with open(dst, "w") as fh:
            fh.write(text)
        if f.endswith(".cu"):
            cpps.append(dst)
    flags = ["-O1", "-g", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-fPIC", "-shared", "-D__CUDACC__",
             "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-function", "-Wno-unused-variable", "-Wno-sign-compare",
             "-I", os.path.join(HERE, "evp"), "-I", os.path.join(BUILD, "csrc")]
    if asan:
        flags += ["-fsanitize=address", "-fsanitize=undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer"]
    # one compiler process per translation unit, then the link
    objs = [c[:-4] + ("_asan.o" if asan else ".o") for c in cpps]
    procs = [subprocess.Popen(["g++"] + [f for f in flags if f != "-shared"] + ["-c", "-o", o, c]) for c, o in zip(cpps, objs)]
    if any(p.wait() != 0 for p in procs):
        raise RuntimeError("the emulated EVP library does not compile")
    subprocess.run(["g++", "-shared"] + (["-fsanitize=address", "-fsanitize=undefined"] if asan else []) + ["-o", out] + objs + ["-ldl"],
                   check=True)
    return out


if __name__ == "__main__":
    print(library())
