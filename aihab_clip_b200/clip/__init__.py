"""Drop-in for the reference ``clip`` package (clip/__init__.py re-exports clip.clip)."""
from .clip import available_models, load, tokenize  # noqa: F401
from . import model  # noqa: F401
