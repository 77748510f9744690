"""Builds the C++ host mirror (include/GpuFMSearcher.hpp) against libfmgpu.so; runs it on the GPU box."""
import os
import subprocess

import pytest

from findex_b200 import build as fbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "findex_b200")


def _compile(tmp_path):
    fbuild.build()
    exe = str(tmp_path / "test_host_mirror")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp"),
                           "-o", exe, os.path.join(LIBDIR, "libfmgpu.so"), "-Wl,-rpath," + LIBDIR])
    return exe


def test_cpp_host_mirror_compiles_and_links(tmp_path):
    _compile(tmp_path)


@pytest.mark.gpu
def test_cpp_host_mirror_runs_reference_assertions(tmp_path, ref_dir):
    exe = _compile(tmp_path)
    out = subprocess.run([exe, ref_dir], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "cpp host mirror ok" in out.stdout
