"""TEST INFRASTRUCTURE — builds ``oracle/_ref/``: the reference's own Python files for the hot path, copied UNMODIFIED
from ``/root/reference`` at build time (the Python counterpart of compiling a C reference into ``oracle/_ref``).

``oracle/_ref/`` is git-ignored (no reference source enters the history) but travels to the GPU box with the
snapshot, so that there (a) ``bench.py --impl reference`` times the REAL reference — ``clip.load`` ->
``build_clip_transforms`` -> ``encode_image`` -> proj / normalise / logits (clip/clip.py:89-137,
data/clip_transforms.py:26-56, methods/utils.py:175-189) — and (b) the drop-in tests feed OUR model to the reference's
own ``compute_image_features`` / ``compute_image_features_test`` / ``clip_classifier`` / cache writers.
Only tests/, __graft_entry__ and bench.py's CPU legs may import it (see oracle/__init__.py).

    python -m oracle.build_ref          (also run by __graft_entry__.build() when /root/reference exists)
"""
from __future__ import annotations

import atexit
import shutil
import sys
import tempfile
import zipfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_SRC = Path("/root/reference")
REF_DST = HERE / "_ref"
ARCHIVE = REF_DST / "reference_path.zip"   # ONE built artefact (like a compiled .so), unpacked to a temp dir on import

# what the path needs, relative to the reference root
FILES = [
    "clip/__init__.py", "clip/clip.py", "clip/model.py", "clip/simple_tokenizer.py", "clip/bpe_simple_vocab_16e6.txt.gz",
    "data/__init__.py", "data/clip_transforms.py", "data/data_utils.py", "data/templates.py",
    "methods/utils.py", "utils.py", "aihab_utils/feature_cache.py",
]
# the only import of the reference's clip package that this image lacks; fix_text is the identity on ASCII prompts
# (clip/simple_tokenizer.py:50-53).  Our stub, not reference code.
FTFY_STUB = '"""stub of the missing `ftfy` dependency (oracle/build_ref.py): identity on ASCII prompts."""\n\n\ndef fix_text(s):\n    return s\n'


def build(verbose: bool = True) -> bool:
    """Packs the files into oracle/_ref/reference_path.zip; returns False (and leaves an existing archive alone) when
    /root/reference is absent (the GPU box: the archive built in the container travels with the snapshot)."""
    if not REF_SRC.is_dir():
        if verbose:
            print(f"[oracle/_ref] {REF_SRC} not present; keeping {'the existing' if ARCHIVE.is_file() else 'no'} archive")
        return False
    if REF_DST.exists():
        shutil.rmtree(REF_DST)
    REF_DST.mkdir(parents=True)
    with zipfile.ZipFile(ARCHIVE, "w", zipfile.ZIP_DEFLATED) as z:
        for rel in FILES:
            z.write(REF_SRC / rel, rel)
        z.writestr("ftfy.py", FTFY_STUB)
        z.writestr("SOURCE.txt", "unmodified copies from /root/reference (WhiteGiveFive/aihab-clip); built by "
                                 "oracle/build_ref.py; git-ignored\n" + "\n".join(FILES) + "\n")
    if verbose:
        print(f"[oracle/_ref] packed {len(FILES)} reference files -> {ARCHIVE} ({ARCHIVE.stat().st_size} bytes)")
    return True


def available() -> bool:
    return ARCHIVE.is_file()


_unpacked = None


def import_ref():
    """Unpacks the archive to a temp dir (the tokenizer opens its vocabulary by file path), puts it at the front of
    sys.path and returns the reference's ``clip`` package.  The reference's top-level names (clip, data, methods,
    utils, aihab_utils) do not collide with this repo's package (everything of ours lives under ``aihab_clip_b200``)."""
    global _unpacked
    if not available():
        raise RuntimeError("oracle/_ref is not built: run `python -m oracle.build_ref` where /root/reference exists")
    if _unpacked is None:
        _unpacked = Path(tempfile.mkdtemp(prefix="aihab_ref_"))
        atexit.register(shutil.rmtree, _unpacked, True)
        with zipfile.ZipFile(ARCHIVE) as z:
            z.extractall(_unpacked)
        sys.path.insert(0, str(_unpacked))
    import clip
    if not Path(clip.__file__).resolve().is_relative_to(_unpacked.resolve()):
        raise RuntimeError(f"`clip` resolved to {clip.__file__}, not to the unpacked oracle/_ref archive")
    return clip


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
