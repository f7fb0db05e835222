"""TEST INFRASTRUCTURE ONLY — imports the *unmodified* reference (/root/reference) in this container.

Nothing under toist_b200/ may import this module.  It exists to (1) validate oracle/model.py (our CPU restatement)
against the reference's own modules and (2) generate the golden fixtures under tests/golden/ (tools/make_golden.py).
/root/reference does not exist on the GPU box, so nothing that runs there may call `load_reference()`.

The five monkey patches follow SURVEY.md Appendix B; the reference sources are never edited or copied.
"""
from __future__ import annotations

import argparse
import sys
from types import ModuleType

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    import os

    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def _install_import_shims() -> None:
    if "IPython" not in sys.modules:
        ip = ModuleType("IPython")
        ip.embed = lambda *a, **k: None
        sys.modules["IPython"] = ip
    import transformers  # noqa: F401  (must precede the fake timm)

    if "timm" not in sys.modules:
        timm = ModuleType("timm")
        timm_models = ModuleType("timm.models")
        timm_models.create_model = None
        timm.models = timm_models
        sys.modules["timm"] = timm
        sys.modules["timm.models"] = timm_models
    import torchvision

    if not getattr(torchvision.models, "_toist_oracle_patched", False):
        for name in ("resnet50", "resnet101"):
            orig = getattr(torchvision.models, name)

            def wrapped(*a, pretrained=False, _orig=orig, **k):
                k.pop("weights", None)  # toist_b200's own builder passes weights=None itself
                return _orig(*a, weights=None, **k)

            setattr(torchvision.models, name, wrapped)
        torchvision.models._toist_oracle_patched = True


def load_reference(tokenizer):
    """Returns the reference's `models` package with the tokenizer factory replaced by `tokenizer`."""
    if not reference_available():
        raise RuntimeError("/root/reference is not present (this is expected on the GPU box)")
    _install_import_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import transformers

    transformers.RobertaTokenizerFast.from_pretrained = classmethod(lambda cls, *a, **k: tokenizer)
    import models  # the reference package

    import models.transformer as ref_tr

    ref_tr.RobertaTokenizerFast.from_pretrained = classmethod(lambda cls, *a, **k: tokenizer)
    return models


def reference_args(extra=()):
    """argparse namespace exactly as the reference's main.py would build it (main.py:32-274,297-298)."""
    _install_import_shims()
    for mod in ("pycocotools", "pycocotools.mask", "pycocotools.coco", "pycocotools.cocoeval"):
        if mod not in sys.modules:
            m = ModuleType(mod)
            m.COCO = object
            m.COCOeval = object
            sys.modules[mod] = m
    sys.modules["pycocotools"].mask = sys.modules["pycocotools.mask"]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import main as ref_main

    parser = argparse.ArgumentParser(parents=[ref_main.get_args_parser()])
    args = parser.parse_args(["--dataset_config", "unused", "--device", "cpu", "--without_pretrain", *extra])
    args.masks = args.mask_model != "none"
    return args
