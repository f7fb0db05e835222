"""`from toist_b200.models import build_model` replaces `from models import build_model` (reference main.py:28,317)."""
from .mdetr import build


def build_model(args):
    """Returns (model, criterion, cluster_criterion | None, weight_dict), as reference models/__init__.py:6-7."""
    return build(args)
