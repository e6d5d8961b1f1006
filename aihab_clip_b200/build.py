"""Builds aihab_clip_b200/libaihab_clip.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m aihab_clip_b200.build [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
TARGET = Path(__file__).resolve().parent / "libaihab_clip.so"


def build(force: bool = False, verbose: bool = True) -> Path:
    if force:
        subprocess.run(["make", "-C", str(CSRC), "clean"], check=True, capture_output=not verbose)
    jobs = str(min(8, os.cpu_count() or 1))
    proc = subprocess.run(["make", "-C", str(CSRC), "-j", jobs], capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc build of libaihab_clip.so failed")
    if verbose:
        print(proc.stdout.strip().splitlines()[-1] if proc.stdout.strip() else "up to date")
    if not TARGET.is_file():
        raise RuntimeError(f"{TARGET} was not produced")
    return TARGET


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
