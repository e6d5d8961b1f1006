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


def gen_prompts(use_hierarchy: bool = True, use_descriptive: bool = False):
    """data/templates.py:236-297 without the descriptive-attribute variants (their attribute dictionaries are
    dataset prose, not part of the hot path): returns (prompts, templates_per_class)."""
    if use_descriptive:
        raise NotImplementedError("descriptive prompt attributes are not carried by the B200 hot-path package")
    prompts = []
    for name, l2 in _L3:
        l3 = name.replace("_", " ")
        if use_hierarchy:
            prompts.append(f"a habitat photo of {L2_NAMES[l2]}, specifically {l3}")
        else:
            prompts.extend(t.format(l3) for t in CS_TEMPLATES)
    return prompts, 1
