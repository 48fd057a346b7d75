"""Drop-in for the reference's ``training/clip`` package (training/clip/__init__.py:1-2)."""
from .clip import available_models, load, tokenize  # noqa: F401
from .model import CLIP, build_model, contrastive_loss  # noqa: F401
