"""CPU-side checks of the drop-in boundary: the in-tree shared library loads and exports
every symbol include/instruct_b200.h declares, the ctypes structs have the C layout, and
without a CUDA device the product path fails loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import instruct_b200
from instruct_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "instruct_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ig_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = instruct_b200.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in instruct_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (ig_[a-z_0-9]+)", out))
    assert set(names) <= exported
    assert lib.ig_version().startswith(b"instruct_b200")


def test_struct_layout_matches_header(tmp_path):
    """sizeof/offsetof of ig_config and ig_chain_result as gcc sees the header."""
    prog = tmp_path / "layout.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "instruct_b200.h"\n'
                    'int main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(ig_config), offsetof(ig_config, alpha_dpm),'
                    'offsetof(ig_config, update), offsetof(ig_config, seed), offsetof(ig_config, device),'
                    'sizeof(ig_chain_result), offsetof(ig_chain_result, indvlkh));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    want = [C.sizeof(_lib.IgConfig), _lib.IgConfig.alpha_dpm.offset, _lib.IgConfig.update.offset,
            _lib.IgConfig.seed.offset, _lib.IgConfig.device.offset, C.sizeof(_lib.IgChainResult),
            _lib.IgChainResult.indvlkh.offset]
    assert got == want


@pytest.mark.skipif(instruct_b200.load().ig_device_count() > 0, reason="a CUDA device is present")
def test_no_gpu_means_loud_failure_not_fallback():
    from instruct_b200.synth import make_dataset
    d = make_dataset(20, 10, 2, A=3, seed=0)
    with pytest.raises(instruct_b200.InstructError, match="no CUDA device"):
        instruct_b200.Sampler(instruct_b200.SeqData(d.x, d.allelenum, 2))
    with pytest.raises(instruct_b200.InstructError):
        instruct_b200.mcmc_updating(instruct_b200.SeqData(d.x, d.allelenum, 2), instruct_b200.Init(10, 5, 1), 0, None)


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under instruct_b200/ may reference it."""
    pkg = os.path.join(ROOT, "instruct_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                for bad in ("pyoracle", "liboracle", "import oracle", "from oracle", "instruct_oracle"):
                    assert bad not in txt, (bad, os.path.join(dirpath, f))
