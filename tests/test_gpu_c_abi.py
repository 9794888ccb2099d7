"""The boundary really is a C ABI: a plain C99 program (tests/c/abi_smoke.c — no torch, no C++) is compiled with gcc
against include/tasu_bridge.h, linked to libtasu_bridge.so and run on the GPU; it checks the PSD decision table of
SURVEY §8(a) through tasu_frame_stats / tasu_collapse_plan / tasu_collapse_scan."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "abi_smoke.c")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _compile(tmp_path, link):
    import ps_slm_b200._lib as L
    out = str(tmp_path / ("abi_smoke" if link else "abi_smoke.o"))
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", SRC, "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(CUDA, "include")]
    if link:
        libdir = os.path.dirname(L.LIB_PATH)
        cmd += ["-L" + libdir, "-ltasu_bridge", "-Wl,-rpath," + libdir, "-L" + os.path.join(CUDA, "lib64"), "-lcudart",
                "-Wl,-rpath," + os.path.join(CUDA, "lib64"), "-o", out]
    else:
        cmd += ["-c", "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


@pytest.mark.skipif(shutil.which("gcc") is None or not os.path.isdir(os.path.join(CUDA, "include")), reason="needs gcc + CUDA headers")
def test_header_is_plain_c99(tmp_path):
    """CPU: the public header and the client compile as C99 with -Wall -Werror."""
    _compile(tmp_path, link=False)


@pytest.mark.gpu
def test_c_client_runs_the_psd_decision_table(tmp_path):
    exe = _compile(tmp_path, link=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "C ABI SMOKE OK" in r.stdout, r.stdout + r.stderr
