"""Pins oracle/clip_oracle.py to the reference: every oracle function is checked against outputs of the
unmodified reference (tests/golden/reference_outputs.npz, produced by tests/golden/make_golden.py)."""
import hashlib

import numpy as np
import pytest

from aihab_clip_b200.weights import GEOMETRIES, make_state_dict_np, state_dict_digest, synthetic_images_u8
from oracle import clip_oracle as O

CASES = {"tiny16": ("ViT-tiny/16", 0, 6, 64), "tiny14": ("ViT-tiny/14", 1, 6, 111), "b32": ("ViT-B/32", 0, 8, 439)}


def case_inputs(tag):
    geom, seed, n, side = CASES[tag]
    sd = make_state_dict_np(geom, seed)
    u8 = np.concatenate([synthetic_images_u8(n // 2, side, seed=1234),
                         synthetic_images_u8(n - n // 2, side, seed=1234, start=n // 2, smooth=True)])
    return GEOMETRIES[geom], sd, u8


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()


@pytest.mark.parametrize("tag", ["tiny16", "tiny14", "b32"])
def test_weights_generator_is_pinned(gold, tag):
    _, sd, _ = case_inputs(tag)
    assert bytes.fromhex(state_dict_digest(sd)) == gold[f"{tag}_digest"].tobytes()


@pytest.mark.parametrize("tag", ["tiny16", "tiny14", "b32"])
def test_encode_image_and_scoring_match_reference(gold, tag):
    geom, sd, u8 = case_inputs(tag)
    R = geom.image_resolution
    x = np.stack([O.clip_preprocess(im, R) for im in u8])
    assert sha(x) == gold[f"{tag}_pre_sha"].tobytes(), "preprocessing must be bit-exact"
    feats = O.encode_image(sd, x)
    np.testing.assert_allclose(feats, gold[f"{tag}_feats"], atol=2e-4, rtol=0)
    emb, logits, top3 = O.score(feats, sd["visual.proj"], gold[f"{tag}_text_w"], 100.0, 3)
    np.testing.assert_allclose(emb, gold[f"{tag}_emb"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(logits, gold[f"{tag}_logits"], atol=5e-4, rtol=0)
    # scoring alone, from the reference's own features: indices must be identical
    _, logits2, top3b = O.score(gold[f"{tag}_feats"], sd["visual.proj"], gold[f"{tag}_text_w"], 100.0, 3)
    ref = gold[f"{tag}_logits"]
    gaps = np.abs(np.diff(np.sort(ref, axis=1)[:, ::-1][:, :4], axis=1)).min(axis=1)
    untied = gaps > 1e-4
    assert (top3b[untied] == gold[f"{tag}_top3"][untied]).all()
    assert (top3b[untied, 0] == gold[f"{tag}_argmax"][untied]).all()


@pytest.mark.parametrize("tag", ["tiny16", "b32"])
def test_layer_trace_matches_reference(gold, tag):
    geom, sd, u8 = case_inputs(tag)
    x = np.stack([O.clip_preprocess(im, geom.image_resolution) for im in u8[:2]])
    _, trace = O.encode_image(sd, x, return_layers=True)
    ref = gold[f"{tag}_trace"]
    got = np.stack(trace)[:, :ref.shape[1], :ref.shape[2]]
    np.testing.assert_allclose(got, ref, atol=3e-4, rtol=0)


@pytest.mark.parametrize("tag", ["tiny16", "tiny14", "b32"])
def test_encode_text_and_text_head(gold, tag):
    _, sd, _ = case_inputs(tag)
    before, emb = O.encode_text(sd, gold[f"{tag}_tok"])
    np.testing.assert_allclose(before, gold[f"{tag}_text_before"], atol=2e-4, rtol=0)
    np.testing.assert_allclose(emb, gold[f"{tag}_text_emb"], atol=2e-4, rtol=0)
    # clip_classifier with one template: column c = normalised embedding of class c's prompt (first 3 classes)
    head = O.text_head([e[None, :] for e in emb])
    np.testing.assert_allclose(head, gold[f"{tag}_text_w"][:, :3], atol=2e-5, rtol=0)
    np.testing.assert_array_equal(gold[f"{tag}_texts"][:3], gold[f"{tag}_tok"])


def test_preprocess_bit_exact_all_cases(gold, meta):
    for (h, w, R) in meta["pre_cases"]:
        rng = np.random.Generator(np.random.PCG64([99, h, w, R]))
        u8 = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
        u8[1] = synthetic_images_u8(1, max(h, w), seed=5, smooth=True)[0][:h, :w]
        y = np.stack([O.clip_preprocess(im, R) for im in u8])
        np.testing.assert_array_equal(y[:, :, ::37, ::41], gold[f"pre_{h}x{w}_{R}_sample"])
        assert sha(y) == gold[f"pre_{h}x{w}_{R}_sha"].tobytes(), (h, w, R)


def test_evaluation_functions(gold, meta):
    logits, labels = gold["ev_logits"], gold["ev_labels"]
    for red in ("sum", "mean", "logsumexp"):
        got = O.aggregate_logits_to_l2(logits, meta["l3_to_l2"], len(meta["l2_names"]), red)
        np.testing.assert_allclose(got, gold[f"ev_l2_{red}"], atol=1e-5, rtol=1e-6)
    np.testing.assert_array_equal(np.asarray(meta["l3_to_l2"])[labels], gold["ev_l2_targets"])
    correct, idx, probs = O.top3_metrics(logits, labels)
    assert correct == int(gold["ev_top3_correct"])
    np.testing.assert_array_equal(idx, gold["ev_top3_idx"])
    np.testing.assert_allclose(probs, gold["ev_top3_probs"], atol=1e-6, rtol=1e-5)
    assert O.cls_acc(logits, labels, 1) == pytest.approx(float(gold["ev_acc1"]))
    assert O.cls_acc(logits, labels, 3) == pytest.approx(float(gold["ev_acc3"]))
    with pytest.raises(ValueError):
        O.aggregate_logits_to_l2(logits[:, :5], meta["l3_to_l2"], 11)
    with pytest.raises(ValueError):
        O.aggregate_logits_to_l2(logits, meta["l3_to_l2"], 11, "median")


def test_text_head_18x80_shape(gold):
    w = gold["b32_text_w_18x80"]
    assert w.shape == (512, 18)
    np.testing.assert_allclose(np.linalg.norm(w, axis=0), 1.0, atol=1e-5)


def test_agreement_set_subset_matches_oracle(gold):
    """First 32 noise + 32 smooth images of the 4096-image agreement set through the oracle."""
    geom = GEOMETRIES["ViT-B/32"]
    sd = make_state_dict_np(geom, 0, with_text=False)
    n = 4096
    u8 = np.concatenate([synthetic_images_u8(16, 224, seed=777),
                         synthetic_images_u8(16, 224, seed=777, start=n // 2, smooth=True)])
    x = np.stack([O.clip_preprocess(im, 224) for im in u8])
    _, logits, _ = O.score(O.encode_image(sd, x), sd["visual.proj"], gold["b32_text_w"], 100.0, 1)
    ref = np.concatenate([gold["agree_logits"][:16], gold["agree_logits"][n // 2:n // 2 + 16]])
    np.testing.assert_allclose(logits, ref, atol=1e-3, rtol=0)


def test_prototype_scores_match_reference_golden():
    """oracle.prototype_scores against tools/outlier_cleaning.py of the unmodified reference
    (tests/golden/make_golden_prototypes.py -> prototype_scores.npz)."""
    from pathlib import Path
    g = np.load(Path(__file__).resolve().parent / "golden" / "prototype_scores.npz")
    sim, pid, other, margin = O.prototype_scores(g["emb"], g["labels"], g["prototypes"], g["owner"])
    np.testing.assert_allclose(sim, g["sim_to_prototype"], atol=2e-6, rtol=0)
    np.testing.assert_array_equal(pid, g["prototype_id"])
    np.testing.assert_allclose(other, g["sim_to_other_class_best"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(margin, g["margin_to_other_class"], atol=4e-6, rtol=0)
    # one prototype per class = the single-centroid scorer
    sim_c, pid_c, _, _ = O.prototype_scores(g["emb"], g["labels"], g["centroids"], g["centroid_owner"])
    np.testing.assert_allclose(sim_c, g["sim_to_centroid"], atol=2e-6, rtol=0)
    assert (pid_c == 0).all()
    # a single class has no "other" class: NaN like the reference
    one = g["labels"] == g["labels"][0]
    _, _, o1, m1 = O.prototype_scores(g["emb"][one], g["labels"][one], g["prototypes"][g["owner"] == g["labels"][0]],
                                      g["owner"][g["owner"] == g["labels"][0]])
    assert np.isnan(o1).all() and np.isnan(m1).all()


@pytest.mark.parametrize("tag", ["tiny16", "tiny14", "b32"])
def test_torch_restatement_matches_reference(gold, tag):
    """oracle/clip_oracle_torch.py (the restatement timed as the CPU reference arm) against the same reference
    goldens: preprocessing bit-exact, features / logits within the numpy oracle's tolerances, identical top-3."""
    import torch
    from oracle import clip_oracle_torch as OT
    geom, sd, u8 = case_inputs(tag)
    x = OT.preprocess_pil(u8, geom.image_resolution)
    assert sha(x.numpy()) == gold[f"{tag}_pre_sha"].tobytes(), "preprocessing must be bit-exact"
    sdt = OT.to_torch_state(sd)
    feats = OT.encode_image(sdt, x)
    np.testing.assert_allclose(feats.numpy(), gold[f"{tag}_feats"], atol=2e-4, rtol=0)
    emb, logits, top3 = OT.score(feats, sdt["visual.proj"], torch.from_numpy(gold[f"{tag}_text_w"]), 100.0, 3)
    np.testing.assert_allclose(emb.numpy(), gold[f"{tag}_emb"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(logits.numpy(), gold[f"{tag}_logits"], atol=5e-4, rtol=0)
    ref = gold[f"{tag}_logits"]
    gaps = np.abs(np.diff(np.sort(ref, axis=1)[:, ::-1][:, :4], axis=1)).min(axis=1)
    untied = gaps > 2e-3
    assert (top3.numpy()[untied] == gold[f"{tag}_top3"][untied]).all()


@pytest.mark.parametrize("tag,geom_name,n", [("lmini", "ViT-L-mini/14", 5), ("lmini336", "ViT-L-mini/14@336px", 3)])
def test_vit_l_geometries_are_pinned_to_the_reference(gold_vitl, tag, geom_name, n):
    """BASELINE.json configs 3 and 4 (ViT-L/14: width 1024, 16 heads, 257 / 577 tokens, 588-wide patches): both
    restatements against outputs of the unmodified reference on the inputs of the GPU parity test
    (tests/test_gpu_model.py::test_vit_l_geometries_match_oracle)."""
    import torch
    from oracle import clip_oracle_torch as OT
    geom = GEOMETRIES[geom_name]
    sd = make_state_dict_np(geom, 3, with_text=False)
    u8 = synthetic_images_u8(n, 300, smooth=True)
    x = np.stack([O.clip_preprocess(im, geom.image_resolution) for im in u8])
    assert sha(x) == gold_vitl[f"{tag}_pre_sha"].tobytes(), "preprocessing must be bit-exact"
    feats = O.encode_image(sd, x)
    np.testing.assert_allclose(feats, gold_vitl[f"{tag}_feats"], atol=3e-4, rtol=0)
    emb, _, _ = O.score(feats, sd["visual.proj"], np.eye(geom.embed_dim, 4, dtype=np.float32), 100.0, 1)
    np.testing.assert_allclose(emb, gold_vitl[f"{tag}_emb"], atol=5e-6, rtol=0)
    sdt = OT.to_torch_state(sd)
    xt = OT.preprocess_pil(u8, geom.image_resolution)
    assert sha(xt.numpy()) == gold_vitl[f"{tag}_pre_sha"].tobytes()
    np.testing.assert_allclose(OT.encode_image(sdt, xt).numpy(), gold_vitl[f"{tag}_feats"], atol=3e-4, rtol=0)


def b16_inputs():
    n, side = 16, 300
    return np.concatenate([synthetic_images_u8(n // 2, side, seed=1234),
                           synthetic_images_u8(n - n // 2, side, seed=1234, start=n // 2, smooth=True)])


def test_headline_geometry_is_pinned_to_the_reference(gold_vitl):
    """BASELINE.json configs[1] — the full ViT-B/16 tower (197 tokens, 12 blocks, width 768) with the shipped 20-class
    text head: both restatements against the unmodified reference's features and x100 logits."""
    import torch
    from oracle import clip_oracle_torch as OT
    geom = GEOMETRIES["ViT-B/16"]
    sd = make_state_dict_np(geom, 0, with_text=False)
    u8 = b16_inputs()
    x = np.stack([O.clip_preprocess(im, geom.image_resolution) for im in u8])
    assert sha(x) == gold_vitl["b16_pre_sha"].tobytes(), "preprocessing must be bit-exact"
    tw = gold_vitl["b16_text_w"]
    emb, logits, _ = O.score(O.encode_image(sd, x), sd["visual.proj"], tw, 100.0, 3)
    np.testing.assert_allclose(emb, gold_vitl["b16_emb"], atol=5e-6, rtol=0)
    np.testing.assert_allclose(logits, gold_vitl["b16_logits"], atol=5e-4, rtol=0)
    sdt = OT.to_torch_state(sd)
    _, logits_t, _ = OT.score(OT.encode_image(sdt, OT.preprocess_pil(u8, geom.image_resolution)), sdt["visual.proj"],
                              torch.from_numpy(tw), 100.0, 3)
    np.testing.assert_allclose(logits_t.numpy(), gold_vitl["b16_logits"], atol=5e-4, rtol=0)


def test_full_depth_goldens_pin_the_torch_restatement(gold, gold_full):
    """The full-depth fixtures of tests/golden/make_golden_full.py (unmodified reference, CPU fp32): the torch-operator
    restatement reproduces the reference's ViT-L/14 features for the first two images of the config-3 set (24 blocks,
    width 1024) and the config-1 logits (ViT-B/32, 18 x 80 head) for the first eight images."""
    import torch
    from oracle import clip_oracle_torch as OT
    geom = GEOMETRIES["ViT-L/14"]
    sdt = OT.to_torch_state(make_state_dict_np(geom, 0, with_text=False))
    u8 = np.concatenate([synthetic_images_u8(4, 300, seed=1234), synthetic_images_u8(4, 300, seed=1234, start=4, smooth=True)])
    x = OT.preprocess_pil(u8, 224)
    assert sha(x.numpy()) == gold_full["l14_pre_sha"].tobytes(), "preprocessing must be bit-exact"
    feats = OT.encode_image(sdt, x[:2])
    np.testing.assert_allclose(feats.numpy(), gold_full["l14_feats"][:2], atol=5e-4, rtol=0)
    geom = GEOMETRIES["ViT-B/32"]
    sdt = OT.to_torch_state(make_state_dict_np(geom, 0, with_text=False))
    u8 = np.concatenate([synthetic_images_u8(32, 224, seed=4321), synthetic_images_u8(32, 224, seed=4321, start=32, smooth=True)])
    _, logits, top3 = OT.score(OT.encode_image(sdt, OT.preprocess_pil(u8[:8], 224)), sdt["visual.proj"],
                               torch.from_numpy(gold["b32_text_w_18x80"]), 100.0, 3)
    np.testing.assert_allclose(logits.numpy(), gold_full["c1_logits"][:8], atol=5e-4, rtol=0)
