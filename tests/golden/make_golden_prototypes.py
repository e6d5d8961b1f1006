"""Golden vectors for the multi-prototype / centroid scoring row, produced by the UNMODIFIED
reference: tools/outlier_cleaning.py MultiPrototypeScorer.score_prototype_distance (:553-668) and
SingleCentroidScorer.compute_centroids / score_centroid_distance (:250-337).

    python tests/golden/make_golden_prototypes.py        (needs /root/reference; writes prototype_scores.npz)
"""
import sys
from pathlib import Path

import numpy as np
import pandas as pd
import torch

sys.path.insert(0, "/root/reference")
from tools.outlier_cleaning import MultiPrototypeResult, MultiPrototypeScorer, SingleCentroidScorer  # noqa: E402

rng = np.random.default_rng(2024)
n, E, classes = 700, 64, [0, 2, 3, 7, 11]          # label ids with gaps, like a filtered class list
labels = rng.choice(classes, size=n)
centers = rng.standard_normal((len(classes), 3, E))
emb = np.stack([centers[classes.index(c), rng.integers(0, 3)] + 0.7 * rng.standard_normal(E) for c in labels])
emb = (emb / np.linalg.norm(emb, axis=1, keepdims=True)).astype(np.float32)
k_per = {0: 1, 2: 3, 3: 2, 7: 4, 11: 2}
protos, counts = {}, {}
for c in classes:
    p = rng.standard_normal((k_per[c], E)).astype(np.float32) + centers[classes.index(c), :1].astype(np.float32)
    protos[c] = torch.from_numpy(p / np.linalg.norm(p, axis=1, keepdims=True))
    counts[c] = [int(v) for v in rng.integers(5, 50, k_per[c])]
meta = pd.DataFrame({"file_name": [f"img_{i:05d}.jpg" for i in range(n)], "ground_truth_num_label": labels})
t_emb, t_lab = torch.from_numpy(emb), torch.from_numpy(labels.astype(np.int64))

scorer = MultiPrototypeScorer(t_emb, t_lab, meta)
res = MultiPrototypeResult(prototypes=protos, class_counts={c: int((labels == c).sum()) for c in classes},
                           prototype_counts=counts, k_per_class=k_per, dim=E)
df = scorer.score_prototype_distance(prototypes=res).sort_values("file_name").reset_index(drop=True)

single = SingleCentroidScorer(t_emb, t_lab, meta)
cen = single.compute_centroids()
dfc = single.score_centroid_distance().sort_values("file_name").reset_index(drop=True)

out = {
    "emb": emb, "labels": labels.astype(np.int64),
    "prototypes": torch.cat([protos[c] for c in sorted(protos)]).numpy(),
    "owner": np.concatenate([[c] * k_per[c] for c in sorted(protos)]).astype(np.int64),
    "sim_to_prototype": df["sim_to_prototype"].to_numpy(np.float32),
    "prototype_id": df["prototype_id"].to_numpy(np.int64),
    "sim_to_other_class_best": df["sim_to_other_class_best"].to_numpy(np.float32),
    "margin_to_other_class": df["margin_to_other_class"].to_numpy(np.float32),
    "centroids": torch.stack([cen.centroids[c] for c in sorted(cen.centroids)]).numpy(),
    "centroid_owner": np.asarray(sorted(cen.centroids), dtype=np.int64),
    "sim_to_centroid": dfc["sim_to_centroid"].to_numpy(np.float32),
}
dst = Path(__file__).resolve().parent / "prototype_scores.npz"
np.savez_compressed(dst, **out)
print(dst, {k: v.shape for k, v in out.items()})
