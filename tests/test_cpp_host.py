"""The C++ host mirrors (include/nlp_lda.hpp, include/nlp_index.hpp) and the reference's tests restated in C++
(tests/cpp/lda_test.cpp, tests/cpp/index_test.cpp): compiled against liblda_b200.so on CPU, executed on the B200 box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "lda_test")


def build(name="lda_test"):
    exe = os.path.join(ROOT, "tests", "cpp", name)
    libdir = os.path.join(ROOT, "nlp_b200")
    if os.path.exists(os.path.join(ROOT, "tests", "cpp", name + ".c")):   # plain C: what cgo compiles the header as
        cmd = ["gcc", "-std=c11", "-O2", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"),
               os.path.join(ROOT, "tests", "cpp", name + ".c"), "-o", exe, "-L", libdir, "-llda_b200", "-lm", "-Wl,-rpath," + libdir]
    else:
        cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"),
               os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-o", exe, "-L", libdir, "-llda_b200", "-Wl,-rpath," + libdir]
    subprocess.check_call(cmd)
    return exe


def test_cpp_host_mirror_compiles_and_links():
    assert os.path.exists(build())
    assert os.path.exists(build("index_test"))
    assert os.path.exists(build("go_shim_sequence_test"))


def test_cpp_host_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([build()], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
    r = subprocess.run([build("index_test")], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_reference_tests_in_cpp():
    r = subprocess.run([build()], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "PASS" in r.stdout


@pytest.mark.gpu
def test_reference_index_tests_in_cpp():
    r = subprocess.run([build("index_test")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "PASS" in r.stdout


@pytest.mark.gpu
def test_go_shim_call_sequence_in_c():
    """go/nlp/lda_b200.go cannot be compiled here (no Go toolchain): its call sequence -- flat-argument entry points,
    malloc'ed pageable buffers, host-side draws for a caller-assigned Rnd, PartialFit, Save / Load -- replayed from plain C
    with lda_test.go's three tests on top"""
    r = subprocess.run([build("go_shim_sequence_test")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "PASS" in r.stdout
