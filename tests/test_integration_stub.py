"""INTEGRATION.md, route B: the pybind11 stub that forwards the reference's four `extension` functions to libschemahead's
host entry points is compiled exactly as printed in the document and checked against the golden vectors of the
reference's own extension (tests/golden/init_apis.npz)."""
import importlib.util
import os
import re
import subprocess
import sys
import sysconfig

import numpy as np
import pytest
import torch

from conftest import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "schemanet-pytorch_b200", "schemanet_b200")
MOD = "sh_integration_stub"        # the document's module is called `extension`; renamed here so that it cannot shadow oracle/_ref


def _build_stub(tmp):
    from torch.utils import cpp_extension as ce
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = re.search(r"```cpp\n(.*?)```", doc, re.S).group(1)
    src = os.path.join(tmp, "extension.cpp")
    open(src, "w").write(code)
    out = os.path.join(tmp, MOD + sysconfig.get_config_var("EXT_SUFFIX"))
    inc = ce.include_paths() + [sysconfig.get_paths()["include"], os.path.join(ROOT, "include")]
    cmd = (["g++", "-std=c++17", "-fPIC", "-shared", "-O1", "-DTORCH_EXTENSION_NAME=" + MOD,
            "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)] + ["-I" + i for i in inc] +
           [src, "-o", out, "-L" + LIBDIR, "-l:libschemahead.so", "-Wl,-rpath," + LIBDIR] +
           ["-L" + p for p in ce.library_paths()] + ["-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_python"])
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    spec = importlib.util.spec_from_file_location(MOD, out)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules.pop(MOD, None)      # keep the process-wide module table clean (the oracle imports the reference's `extension`)
    return mod


@pytest.mark.gpu
def test_pybind_stub_from_integration_md(tmp_path):
    ext = _build_stub(str(tmp_path))
    g = load_golden("init_apis")
    B, M, K, Vc = g["cfg"].tolist()
    t = lambda k: torch.from_numpy(np.ascontiguousarray(g[k]))
    ing, attn, attn_cls, geo = t("ingredients"), t("attn"), t("attn_cls"), t("geo_sim")
    assert np.array_equal(ext.feat_to_v_attr(ing, attn_cls, M, True, False).numpy(), g["v_attr_mean"])
    assert np.array_equal(ext.feat_to_v_attr(ing, attn_cls, M, False, False).numpy(), g["v_attr_sum"])
    dicts = [{int(k): v for v, k in enumerate(row)} for row in g["class_ingredients"]]
    assert np.array_equal(ext.feat_to_e(ing, attn, geo, dicts, g["label"].tolist(), Vc, True).numpy(), g["e_mean"])
    w = torch.tensor([[0.3], [0.7]])
    cat_ids, cat_w, nv = ext.feat_to_instance_v(ing, attn_cls, w, False)
    assert np.array_equal(cat_ids.numpy(), g["iv_sum_ids"]) and np.array_equal(nv.numpy(), g["iv_sum_nv"])
    err = np.abs(cat_w.double().numpy() - g["iv_sum_w"]).max() / np.abs(g["iv_sum_w"]).max()
    assert err <= 1e-5, err
    ids = list(torch.split_with_sizes(cat_ids, nv.tolist()))
    idicts = [{v: k for k, v in enumerate(i.tolist())} for i in ids]
    es = ext.feat_to_instance_e(ing, attn, geo, idicts, w, False, False)
    cat_e = torch.cat([e.reshape(-1) for e in es]).double().numpy()
    err = np.abs(cat_e - g["ie_sum_cat"]).max() / np.abs(g["ie_sum_cat"]).max()
    assert err <= 1e-5, err


def test_pybind_stub_compiles(tmp_path):
    """No GPU needed: the stub printed in INTEGRATION.md compiles and links against torch and libschemahead."""
    if not os.path.exists(os.path.join(LIBDIR, "libschemahead.so")):
        pytest.skip("libschemahead.so not built")
    _build_stub(str(tmp_path))
