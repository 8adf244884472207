"""The C ABI driven from plain C (tests/cabi/test_cabi.c, gcc -std=c99 -pedantic against include/gpr_sm100a.h only).
CPU box: the header compiles as C99, the program links against libgpr_sm100a.so and gpr_ctx_create fails loudly with
GPR_ERR_CUDA (no CPU fallback) -> exit code 77.  GPU box (-m gpu): the known-answer checks of
/root/reference/test/test_loss.jl:1-11 (diagonal covariance) pass through create -> nlml_grad -> update_cache ->
predict -> destroy, including releasing the context before its models."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gaussianprocessregression.jl_b200")
SRC = os.path.join(ROOT, "tests", "cabi", "test_cabi.c")


def _build(tmp_path, built_lib):
    exe = str(tmp_path / "test_cabi")
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-L", PKG,
           "-lgpr_sm100a", "-lm", f"-Wl,-rpath,{PKG}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_self_contained_c99(tmp_path):
    src = tmp_path / "hdr.c"
    src.write_text('#include "gpr_sm100a.h"\nint main(void) { return GPR_T_COUNT > 0 ? 0 : 1; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_c_caller_links_and_fails_loudly_without_gpu(tmp_path, built_lib):
    exe = _build(tmp_path, built_lib)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    import gpr_sm100a._ffi as ffi
    if ffi.device_count() == 0:
        assert r.returncode == 77, (r.returncode, r.stdout, r.stderr)
        assert "no CPU fallback" in r.stdout
    else:
        assert r.returncode == 0, (r.stdout, r.stderr)


@pytest.mark.gpu
def test_c_caller_known_answers_on_gpu(tmp_path, built_lib):
    exe = _build(tmp_path, built_lib)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "test_cabi ok" in r.stdout
