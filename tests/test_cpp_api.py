"""The C ABI driven from compiled code: tests/cpp/reference_specs.cpp uses include/thylacine_b200.hpp (the header-only
C++ mirror of the reference's operator surface) to run the reference's own component specs.  Built with g++ against the
in-tree shared library; nothing here goes through Python's ctypes binding."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "thylacine_b200" / "lib"


@pytest.fixture(scope="module")
def binary(tmp_path_factory):
    out = tmp_path_factory.mktemp("cpp") / "reference_specs"
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "tests" / "cpp" / "reference_specs.cpp"),
           "-o", str(out), f"-L{LIB}", "-lthylacine_b200", f"-Wl,-rpath,{LIB}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return out


def test_cpp_mirror_compiles_and_fails_loudly_without_a_gpu(binary):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the no-device path cannot be exercised")
    r = subprocess.run([str(binary), "--no-gpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "cpp api ok (no gpu)" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_reference_specs_from_cpp(binary):
    r = subprocess.run([str(binary)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "cpp api ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_c_header_is_plain_c99_and_cpp_mirror_is_warning_clean(tmp_path):
    """include/thylacine_b200.h must be consumable by a C compiler (the FFI boundary: plain pointers and sizes, no C++),
    and the C++ mirror must build warning-free under -Wall -Wextra -Wpedantic -Wshadow."""
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Wpedantic", "-Werror", "-fsyntax-only", "-x", "c",
                        str(ROOT / "include" / "thylacine_b200.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    src = tmp_path / "tu.cpp"
    src.write_text('#include "thylacine_b200.hpp"\nint main() { return thylacine::Matrix::identity(2).rows == 2 ? 0 : 1; }\n')
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Wpedantic", "-Wshadow", "-Werror", f"-I{ROOT / 'include'}",
                        "-fsyntax-only", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
