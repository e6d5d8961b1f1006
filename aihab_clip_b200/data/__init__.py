"""Label maps and prompt templates the text head is built from (data constants of the reference:
data/__init__.py:28-49,98-133 and data/templates.py:204-226).  The reference treats these as plain data; they
are restated here so the drop-in builds the same 20-class zero-shot head.  Checked against the reference in
tests/test_host_logic.py via tests/golden/reference_meta.json."""

# L3 habitat classes in label order 0..19 (data/__init__.py:28-49) with their L2 parent id (:112-133)
_L3 = [
    ("Urban", 0), ("Broadleaved Mixed and Yew Woodland", 1), ("Coniferous Woodland", 1), ("Sea", 9),
    ("Arable and Horticulture", 2), ("Improved Grassland", 3), ("Neutral Grassland", 3), ("Calcareous Grassland", 3),
    ("Acid Grassland", 3), ("Bracken", 3), ("Dwarf Shrub Heath", 4), ("Fen, Marsh, Swamp", 5), ("Bog", 5),
    ("Littoral Rock", 6), ("Littoral Sediment", 6), ("Montane", 10), ("Standing Open Waters and Canals", 8),
    ("Inland Rock", 7), ("Supra-littoral Rock", 7), ("Supra-littoral Sediment", 7),
]
# L2 classes in id order (data/__init__.py:98-110)
L2_NAMES = ["Urban", "Woodland and Forest", "Cropland", "Grassland", "Heathland and Shrub", "Wetland",
            "Marine Inlets and Transitional Waters", "Sparsely Vegetated Land", "Rivers and Lakes", "Sea", "Montane"]

REASSIGN_LABEL_NAME_L3 = {i: name for i, (name, _) in enumerate(_L3)}
NAME_LABEL_L2 = {name: i for i, name in enumerate(L2_NAMES)}
REASSIGN_NAME_LABEL_L3L2 = {name: (i, l2) for i, (name, l2) in enumerate(_L3)}

CS_CLASSNAMES = [name for name, _ in _L3]          # data/templates.py:226
CS_TEMPLATES = ["a habitat photo of {}."]           # the one active template, data/templates.py:204-223


def build_l3_to_l2_map():
    """data/__init__.py:253-268 — (l3_to_l2 list indexed by L3 id, L2 names indexed by L2 id)."""
    return [l2 for _, l2 in _L3], list(L2_NAMES)


def _descriptive():
    """Per-class attribute phrases and the two descriptive templates of data/templates.py:12-201 (dataset prose kept as
    a data file, exported from the reference by tools/export_prompt_data.py)."""
    import json
    from pathlib import Path
    return json.loads((Path(__file__).resolve().parent / "descriptive_attrs.json").read_text())


def gen_prompts(use_hierarchy: bool = True, use_descriptive: bool = True):
    """data/templates.py:236-297 — prompts for the 20 L3 classes, class-major: with the L2 context
    (``use_hierarchy``) or the flat CS_TEMPLATES, with the per-class descriptive attributes appended
    (``use_descriptive``; prints a two-prompt preview per class like the reference).  Returns
    ``(prompts, templates_per_class)``."""
    base_templates = ["a habitat photo of {l2}, specifically {l3}"] if use_hierarchy else CS_TEMPLATES
    desc = _descriptive() if use_descriptive else None
    desc_templates = (desc["HIER_DESC_TEMPLATES"] if use_hierarchy else desc["DESC_TEMPLATES"]) if desc else []
    if use_descriptive and len(base_templates) != len(desc_templates):
        raise ValueError(
            "Descriptive templates enabled but template counts differ: "
            f"{len(desc_templates)} (descriptive) vs {len(base_templates)} (base). "
            "Please make them consistent.")
    templates_per_class = len(desc_templates) if use_descriptive else len(base_templates)
    prompts = []
    for name, l2_id in _L3:
        l3 = name.replace("_", " ")
        l2 = L2_NAMES[l2_id]
        attrs = desc["DESCRIPTIVE_L3_ATTRS"].get(l3) if desc else None
        if attrs is not None:
            text = ", ".join(attrs.values())
            cls = [t.format(l2=l2, l3=l3, attrs=text) if use_hierarchy else t.format(habitat=l3, attrs=text)
                   for t in desc_templates]
        elif use_hierarchy:
            cls = [t.format(l3=l3, l2=l2) for t in base_templates]
        else:
            cls = [t.format(l3) for t in base_templates]
        if use_descriptive:
            print(f"[gen_prompts] {l3}: {cls[:min(2, len(cls))]}")
        prompts.extend(cls)
    return prompts, templates_per_class
