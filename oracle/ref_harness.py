"""Import the UNMODIFIED reference head in this container (test infrastructure only).

Only usable where /root/reference exists (the build container, never the GPU box).  Two shims, neither
touching arithmetic:
  1. `cpp_extension` is the reference's own package __init__ (read in place) with its native
     `extension` module resolved from oracle/_ref/ (built by oracle/build_ref.py from the reference sources).
  2. the un-vendored `cv_lib` is replaced by MagicMock modules (import-time dependency only).
"""
import importlib.util
import os
import sys
from unittest import mock

REF = os.environ.get("SCHEMANET_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

_CV_LIB = ["cv_lib", "cv_lib.utils", "cv_lib.config_parsing", "cv_lib.optimizers", "cv_lib.schedulers",
           "cv_lib.distributed", "cv_lib.distributed.utils", "cv_lib.distributed.sampler", "cv_lib.metrics",
           "cv_lib.logger", "cv_lib.classification", "cv_lib.classification.data",
           "cv_lib.classification.models", "cv_lib.augmentation", "cv_lib.utils.cuda_utils",
           "cv_lib.classification.data.imagenet", "cv_lib.classification.data.classification_dataset"]


def available():
    return os.path.isdir(os.path.join(REF, "schema_inference"))


def import_reference():
    """Returns a namespace with the reference's Discretization, SchemaNet, Matcher, GNN, ext functions."""
    if not available():
        raise RuntimeError("reference tree not present: golden vectors can only be (re)generated in the build container")
    sys.path.insert(0, HERE)
    import build_ref
    build_ref.build()
    for name in _CV_LIB:
        sys.modules.setdefault(name, mock.MagicMock())
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import torch  # noqa: F401
    if "cpp_extension" not in sys.modules:
        spec = importlib.util.spec_from_file_location(
            "cpp_extension", os.path.join(REF, "cpp_extension", "__init__.py"),
            submodule_search_locations=[os.path.join(HERE, "_ref")])
        mod = importlib.util.module_from_spec(spec)
        sys.modules["cpp_extension"] = mod
        spec.loader.exec_module(mod)
    import cpp_extension
    from discretization.discretization import Discretization
    from discretization.visual_word_encoder import Adapter
    from schema_inference.graph.schema_net import SchemaNet
    from schema_inference.graph.match import Matcher
    from schema_inference.graph.gnn import GNN
    import schema_inference.graph.utils as graph_utils
    from schema_inference.utils.ingredient_model_wrapper import IngredientModelWrapper

    class NS:
        pass
    ns = NS()
    ns.cpp_extension = cpp_extension
    ns.Discretization = Discretization
    ns.Adapter = Adapter
    ns.SchemaNet = SchemaNet
    ns.Matcher = Matcher
    ns.GNN = GNN
    ns.graph_utils = graph_utils
    ns.IngredientModelWrapper = IngredientModelWrapper
    return ns
