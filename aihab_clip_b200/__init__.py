"""aihab_clip_b200 — B200-native (sm_100a) implementation of the aihab-clip image-encode + zero-shot scoring hot
path behind the reference's Python API.  See DESIGN.md."""
__version__ = "0.1.0"
