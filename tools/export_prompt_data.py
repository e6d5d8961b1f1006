"""Exports the reference's descriptive prompt attributes (data/templates.py:12-201: per-class attribute phrases and the
two descriptive templates — dataset prose, plain data) to aihab_clip_b200/data/descriptive_attrs.json, and the outputs
of the reference's gen_prompts for all four flag combinations to tests/golden/reference_prompts.json.
Run in the build container only (needs /root/reference):   python tools/export_prompt_data.py"""
import contextlib
import io
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests" / "golden"))
from make_golden import import_reference  # noqa: E402

import_reference()
from data import templates as T  # noqa: E402  (the reference module)

data = {"DESC_TEMPLATES": T.DESC_TEMPLATES, "HIER_DESC_TEMPLATES": T.HIER_DESC_TEMPLATES,
        "DESCRIPTIVE_L3_ATTRS": T.DESCRIPTIVE_L3_ATTRS}
(REPO / "aihab_clip_b200" / "data" / "descriptive_attrs.json").write_text(json.dumps(data, indent=1))
gold = {}
for h in (True, False):
    for d in (True, False):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            prompts, tpc = T.gen_prompts(use_hierarchy=h, use_descriptive=d)
        gold[f"hier={h},desc={d}"] = {"prompts": prompts, "templates_per_class": tpc, "stdout": buf.getvalue()}
(REPO / "tests" / "golden" / "reference_prompts.json").write_text(json.dumps(gold, indent=1))
print("wrote descriptive_attrs.json and reference_prompts.json;", {k: len(v["prompts"]) for k, v in gold.items()})
