"""CPU suite, part 2: the drop-in boundary.  The shared library loads, exports every symbol
include/dy4_b200.h declares and every C++ function of the reference's filter.h, the host-side
pieces (mode table, tap design) agree with the oracle, and — without a GPU — every compute
entry point fails loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol(dy4):
    hdr = open(os.path.join(ROOT, "include", "dy4_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(dy4_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 30
    L = C.CDLL(dy4.LIB_PATH)
    missing = [n for n in declared if not hasattr(L, n)]
    assert not missing, missing
    # and the ctypes binding covers exactly the header
    assert sorted(dy4._lib.EXPORTS) == declared


def test_library_exports_reference_cpp_signatures(dy4):
    """What project.cpp needs at link time: the 15 functions of the reference's include/filter.h:17-34."""
    out = subprocess.run("nm -D --defined-only %s | c++filt" % dy4.LIB_PATH, shell=True, capture_output=True, text=True).stdout
    v = "std::vector<float, std::allocator<float> >"
    want = [
        "impulseResponseLPF(float, float, unsigned short, %s&, int)" % v,
        "convolveFIR(%s&, %s const&, %s const&)" % (v, v, v),
        "blockConvolveFIR(%s&, %s const&, %s const&, %s&)" % (v, v, v, v),
        "fmDemodArctan(%s const&, %s const&, float&, float&, %s&)" % (v, v, v),
        "downsample(%s, unsigned long, %s&)" % (v, v),
        "upsample(%s, unsigned long, %s&)" % (v, v),
        "downsampleBlockConvolveFIR(int, %s&, %s const&, %s const&, %s&)" % (v, v, v, v),
        "resampleBlockConvolveFIR(int, int, %s&, %s const&, %s const&, %s&)" % (v, v, v, v),
        "impulseResponseBPF(float, float, float, unsigned short, %s&, int)" % v,
        "fmPLL(%s const&, float, float, float, float, float, %s&, float&, float&, float&, float&, float&, float&)" % (v, v),
        "delayBlock(%s const&, %s&, %s&)" % (v, v, v),
        "pointwiseMultiply(%s const&, %s const&, %s&)" % (v, v, v),
        "pointwiseAdd(%s const&, %s const&, %s&)" % (v, v, v),
        "pointwiseSubtract(%s const&, %s const&, %s&)" % (v, v, v),
        "interleave(%s const&, %s const&, %s&)" % (v, v, v),
    ]
    for w in want:
        assert w in out, w


def test_library_is_sm100a_cuda(dy4):
    out = subprocess.run(["cuobjdump", "-lelf", dy4.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out.stdout


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_mode_table_matches_oracle(dy4, orc, mode):
    a, b = dy4.mode_params(mode), orc.mode_params(mode)
    for f in a._fields:
        assert getattr(a, f) == getattr(b, f), f
    assert a.block_size == 2 * a.rf_decim * a.if_per_block
    assert a.if_per_block * a.audio_upsample == a.audio_per_block * a.audio_decim


def test_bad_arguments_are_rejected(dy4):
    with pytest.raises(dy4.Dy4Error):
        dy4.mode_params(4)
    with pytest.raises(dy4.Dy4Error):
        dy4.Pipeline(7, True, 4)
    with pytest.raises(dy4.Dy4Error):
        dy4.Pipeline(0, True, 0)


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_without_gpu(dy4):
    with pytest.raises(dy4.Dy4Error) as e:
        dy4.Pipeline(0, True, 4)
    assert "(-2)" in str(e.value)                      # DY4_ERR_CUDA
    x = np.zeros(1000, np.float32)
    with pytest.raises(dy4.Dy4Error):
        dy4.filterh.blockConvolveFIR(x, np.ones(101, np.float32), np.zeros(100, np.float32))
    with pytest.raises(dy4.Dy4Error):
        dy4.filterh.fmPLL(x, 19e3, 240e3, 2.0, 0.0, 0.01, np.array([1, 0, 0, 0, 0, 1], np.float32))
    assert dy4.launch_count() == 0


def test_product_does_not_import_the_oracle():
    pk = os.path.join(ROOT, "3dy4-real-time-software-defined-radio-_b200")
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "dy4_oracle" not in src and "libdy4ref" not in src, f


def test_streaming_command_line_argument_handling(dy4):
    """dy4_project (the reference's `project <mode> <mono|stereo>` boundary): usage and argument errors need no GPU."""
    import os
    import subprocess
    exe = os.path.join(dy4.PACKAGE_DIR, "dy4_project")
    assert os.path.exists(exe), "dy4_project not built (make -C csrc)"
    p = subprocess.run([exe], capture_output=True, timeout=60)
    assert p.returncode == 1 and b"Usage" in p.stderr
    p = subprocess.run([exe, "9", "mono"], capture_output=True, timeout=60)
    assert p.returncode == 1 and b"Wrong mode" in p.stderr
    p = subprocess.run([exe, "0", "quad"], capture_output=True, timeout=60)
    assert p.returncode == 1 and b"must be mono or stereo" in p.stderr


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU behaviour")
def test_streaming_command_line_has_no_cpu_fallback(dy4):
    import os
    import subprocess
    exe = os.path.join(dy4.PACKAGE_DIR, "dy4_project")
    p = subprocess.run([exe, "0", "stereo"], input=b"\x80" * 4096, capture_output=True, timeout=120)
    assert p.returncode == 2 and p.stdout == b"" and b"dy4_project:" in p.stderr     # refuses to run: no CUDA device


def test_cmake_file_builds_the_same_library_for_sm_100a():
    """north_star: 'CMakeLists builds for sm_100a with no Triton, no multi-backend dispatch and no CPU fallback'.  The file names
    exactly the sources csrc/Makefile compiles and one architecture."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    txt = open(os.path.join(root, "CMakeLists.txt")).read()
    assert "CMAKE_CUDA_ARCHITECTURES 100a" in txt and len(re.findall(r"CUDA_ARCHITECTURES", txt)) == 1
    mk = open(os.path.join(root, "3dy4-real-time-software-defined-radio-_b200", "csrc", "Makefile")).read()
    cu = re.search(r"^CU := (.*)$", mk, re.M).group(1).split()
    for name in cu:
        assert "/%s.cu" % name in txt, name
    for name in ("dy4_taps.cpp", "filter_shim.cpp", "dy4_project.cpp"):
        assert name in txt
