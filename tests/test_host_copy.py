"""Host-side staging logic of the C-ABI library on the CPU: the streaming-store piece copy and the copy pool (zstdlite_b200/csrc/zl_host.h)
that move pageable / scattered host buffers into and out of pinned staging (what the reference's C layer hands over: src/raw-file.c:166,189)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cuda_include():
    for p in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if p and os.path.exists(os.path.join(p, "include", "cuda_runtime.h")):
            return os.path.join(p, "include")
    nvcc = shutil.which("nvcc")
    if nvcc:
        inc = os.path.join(os.path.dirname(os.path.dirname(os.path.realpath(nvcc))), "include")
        if os.path.exists(os.path.join(inc, "cuda_runtime.h")):
            return inc
    return None


def test_copy_pool_and_streaming_copy(tmp_path):
    inc = _cuda_include()
    if inc is None:
        pytest.skip("cuda_runtime.h not found (the header includes it)")
    exe = str(tmp_path / "copy_pool_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-w", "-I", os.path.join(ROOT, "zstdlite_b200", "csrc"), "-I", inc,
                           os.path.join(ROOT, "tests", "host", "copy_pool_check.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "bad=0" in out.stdout, out.stdout + out.stderr
    assert "threads 5" in out.stdout
