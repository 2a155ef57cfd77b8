"""The C ABI from a C99 caller: include/elph_b200.h must compile as plain C (-std=c99 -pedantic -Werror), every declared
symbol must link against libelph_b200.so, and (on a GPU) a compiled host program drives the operators and a solve."""
import os
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "c_abi" / "abi_check.c"


def _build(tmp_path):
    from elphdynamics_b200 import _lib, build
    lib = build.build()
    syms = _lib.declared_symbols()
    assert len(syms) > 60
    table = ", ".join(f"(void*){s}" for s in syms)
    exe = tmp_path / "abi_check"
    gcc = shutil.which("gcc")
    assert gcc, "gcc is part of the image"
    cmd = [gcc, "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-Wno-pedantic", f"-DSYMBOL_TABLE={table}", f"-I{ROOT / 'include'}",
           str(SRC), "-o", str(exe), f"-L{lib.parent}", "-lelph_b200", "-lm", f"-Wl,-rpath,{lib.parent}"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return exe, len(syms)


def test_header_is_c99_and_every_symbol_links(tmp_path):
    exe, nsym = _build(tmp_path)
    r = subprocess.run([str(exe)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0, r.stdout
    assert f"{nsym} symbols linked" in r.stdout


@pytest.mark.gpu
def test_c_caller_drives_operators_and_solve(tmp_path):
    exe, _ = _build(tmp_path)
    r = subprocess.run([str(exe), "gpu"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    assert "abi_check gpu ok" in r.stdout
