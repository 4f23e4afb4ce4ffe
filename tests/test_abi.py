"""CPU: the C-ABI library loads and exports every symbol include/schemahead.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "schemahead.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sh_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_the_binding_lists():
    from schemanet_b200 import native
    assert _header_symbols() == sorted(native.SYMBOLS)


def test_library_exports_every_declared_symbol():
    from schemanet_b200 import native
    if not os.path.exists(native.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(native.LIB_PATH)
    for name in _header_symbols():
        assert hasattr(lib, name), f"{name} declared in schemahead.h but not exported"
    assert lib.sh_abi_version() == native.ABI_VERSION


def test_built_for_sm100a_only():
    from schemanet_b200 import native
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", native.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback():
    """CUDA-only entry points must refuse CPU tensors instead of silently computing on the host."""
    import torch
    from schemanet_b200 import native
    with pytest.raises(RuntimeError):
        native.discretize(torch.zeros(4, 8), torch.zeros(2, 8))
    with pytest.raises(RuntimeError):
        native.class_atlas(torch.zeros(2, 4), torch.zeros(2, 4, 4))


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "schemanet-pytorch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "head_oracle" not in src and "graph_oracle" not in src and "oracle/" not in src, f
