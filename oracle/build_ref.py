#!/usr/bin/env python
"""Compile the REFERENCE's own cpp_extension into oracle/_ref/ (test infrastructure only).

The sources are compiled where they lie under /root/reference/cpp_extension/src with plain g++
(the reference's setup.py is NOT run: it pins -std=c++14, which torch>=2.1 headers reject;
SURVEY.md section 8c).  Nothing is copied: only the built `extension*.so` lands in oracle/_ref/,
which is git-ignored (but not gpurun-ignored, so it travels to the GPU box).

Usage: python oracle/build_ref.py [--force]
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SCHEMANET_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
# the six translation units listed in the reference's cpp_extension/setup.py:10-17
SOURCES = ["extension.cpp", "feat_to_v_attr.cpp", "feat_to_e.cpp",
           "large_scale_feat_to_v.cpp", "large_scale_feat_to_e.cpp", "utils.cpp"]


def so_path():
    return os.path.join(OUT, "extension" + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force=False):
    src_dir = os.path.join(REF, "cpp_extension", "src")
    if not os.path.isdir(src_dir):
        return None                       # GPU box: use whatever was prebuilt
    target = so_path()
    if os.path.exists(target) and not force:
        return target
    import torch
    from torch.utils import cpp_extension as tce
    os.makedirs(os.path.join(OUT, "obj"), exist_ok=True)
    inc = [os.path.join(REF, "cpp_extension", "include")] + tce.include_paths() + [sysconfig.get_paths()["include"]]
    cflags = ["-O2", "-fPIC", "-std=c++17", "-w", "-DTORCH_EXTENSION_NAME=extension",
              "-DTORCH_API_INCLUDE_EXTENSION_H",
              "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    cflags += ["-I" + p for p in inc]

    def cc(name):
        obj = os.path.join(OUT, "obj", name.replace(".cpp", ".o"))
        subprocess.check_call(["g++", "-c", os.path.join(src_dir, name), "-o", obj] + cflags)
        return obj

    with ThreadPoolExecutor(max_workers=6) as ex:
        objs = list(ex.map(cc, SOURCES))
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    subprocess.check_call(["g++", "-shared", "-o", target] + objs +
                          ["-L" + libdir, "-Wl,-rpath," + libdir,
                           "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python"])
    return target


def load():
    """Import the built reference extension as a module (torch must be imported first)."""
    import importlib.util
    import torch  # noqa: F401  (registers libtorch symbols)
    p = so_path()
    if not os.path.exists(p):
        return None
    spec = importlib.util.spec_from_file_location("extension", p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    t = build(force="--force" in sys.argv)
    print("oracle/_ref:", t)
