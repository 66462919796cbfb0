"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/gpba.h declares; without a CUDA device the product path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from pygpba import lib as gl
from pygpba import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "gpba.h")).read()
    return sorted(set(re.findall(r"^\s*(?:int|void\*?|const char\*)\s+(gpba_\w+)\s*\(", src, re.M)))


def test_library_exports_every_declared_symbol():
    L = gl.lib()
    declared = header_symbols()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/gpba.h but not exported by libgpba.so"
    assert sorted(gl.SYMBOLS) == declared


def test_struct_layouts_match_header():
    """ctypes mirrors must have the sizes the C compiler gives the header structs."""
    import subprocess, tempfile, textwrap
    code = textwrap.dedent("""
        #include <stdio.h>
        #include "gpba.h"
        #include "gpba_map.h"
        int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(gpba_problem), sizeof(gpba_lm_trace), sizeof(gpba_lm_params),
                                sizeof(gpba_thresholds), sizeof(gpba_structure_info), sizeof(gpba_pose_batch), sizeof(gpba_vel_batch),
                                sizeof(gpba_create_options), sizeof(gpba_map_config)); return 0; }
    """)
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(code)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        sizes = list(map(int, subprocess.check_output([os.path.join(d, "t")]).split()))
    from pygpba.problem import CProblem, LmTrace, LmParams, Thresholds, StructureInfo, CreateOptions
    from pygpba.pose import CPoseBatch
    from pygpba.velransac import CVelBatch
    from pygpba.mapmirror import CMapConfig
    assert sizes == [C.sizeof(CProblem), C.sizeof(LmTrace), C.sizeof(LmParams), C.sizeof(Thresholds), C.sizeof(StructureInfo),
                     C.sizeof(CPoseBatch), C.sizeof(CVelBatch), C.sizeof(CreateOptions), C.sizeof(CMapConfig)]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    P = synth.make_problem("tiny")
    with pytest.raises(gl.GpbaError) as e:
        gl.GpBa(P)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_reference_oracle():
    """The product tree must not import / link / execute anything under oracle/."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "amc-slam_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"oracle_py|libgpba_oracle|oracle/|numpy_mirror", txt):
                    bad.append(os.path.join(base, f))
    # the boundary headers, the reference-side bindings and the example neither include nor link anything of the oracle, the
    # compiled reference (oracle/_ref) or the C-ABI test double (oracle/abi_double.cc)
    for d in ("include", "adapter", "examples"):
        for base, _, files in os.walk(os.path.join(ROOT, d)):
            for f in files:
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r'#include\s*[<"][^>"]*(oracle|ref_shim|_ref)|gpba_oracle|abi_double|oracle_[a-z_]+\(', txt):
                    bad.append(os.path.join(base, f))
    assert not bad, bad
    # and the shipped library has no dependency on them
    import subprocess
    so = os.path.join(ROOT, "amc-slam_b200", "libgpba.so")
    if os.path.exists(so):
        needed = subprocess.run(["readelf", "-d", so], capture_output=True, text=True).stdout
        assert "oracle" not in needed and "abi_double" not in needed and "_ref" not in needed


def _build_c_example(tmpdir):
    import subprocess
    exe = os.path.join(tmpdir, "gpba_c_example")
    libdir = os.path.join(ROOT, "amc-slam_b200")
    gl.lib()   # make sure libgpba.so exists
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "gpba_c_example.c"), "-L", libdir, "-lgpba", f"-Wl,-rpath,{libdir}", "-lm", "-o", exe])
    return exe


def test_c_example_links_and_fails_loudly_without_a_device(tmp_path):
    """The boundary is usable from plain C (no torch, no Python): examples/gpba_c_example.c compiles with -Wall -Werror
    against include/gpba.h, links libgpba.so, and on a box without a GPU reports GPBA_ERR_NO_DEVICE (exit code 3)."""
    import subprocess
    import torch
    exe = _build_c_example(str(tmp_path))
    if torch.cuda.is_available():
        pytest.skip("device present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_c_example_optimizes_on_the_gpu(tmp_path):
    import subprocess
    exe = _build_c_example(str(tmp_path))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "LM iterations" in r.stdout
