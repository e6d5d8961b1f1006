#!/bin/bash
# Builds oracle/_ref (git-ignored): unmodified copies of the reference's hot-path files, so that the reference arm of
# bench.py and the drop-in tests run the REAL reference on the GPU box.  See oracle/build_ref.py.
cd "$(dirname "$0")/.." && exec python -m oracle.build_ref
