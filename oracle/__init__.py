"""TEST INFRASTRUCTURE ONLY.  CPU restatements of the reference's algorithm for the hot path (clip_oracle.py: numpy,
clip_oracle_torch.py: the torch CPU operators the reference itself calls) and the recipe that packs the reference's own
files into the git-ignored oracle/_ref/ (build_ref.py).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs
(`cpu_baseline`, `--impl reference`) may import anything from here; the product package aihab_clip_b200 never does, and
fails loudly without its CUDA extension."""
