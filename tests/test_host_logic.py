"""CPU tests of the host side: tokenizer, checkpoint loading protocol, text tower, cache naming / file formats,
evaluation mirrors, sharding arithmetic and the C-ABI export list.  No CUDA calls."""
import ctypes
import json
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from aihab_clip_b200 import _lib, evaluation as EV, feature_cache as FC
from aihab_clip_b200.data import CS_CLASSNAMES, CS_TEMPLATES, build_l3_to_l2_map, gen_prompts
from aihab_clip_b200.extraction import shard_range
from aihab_clip_b200.weights import GEOMETRIES, make_state_dict

REPO = Path(__file__).resolve().parent.parent


def test_c_abi_exports_every_declared_symbol():
    header = (REPO / "include" / "aihab_clip.h").read_text()
    declared = set(re.findall(r"AIHAB_API[^;]*?\b(aihab_\w+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), "ctypes signatures and header drifted apart"
    lib = ctypes.CDLL(str(_lib.LIB_PATH))   # built by __graft_entry__.build(); loading needs no GPU
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().aihab_abi_version() == 1
    assert _lib.kernel_launches() == 0


def test_label_data_matches_reference(meta):
    assert CS_CLASSNAMES == meta["classnames"] and CS_TEMPLATES == meta["templates"]
    l3_to_l2, l2_names = build_l3_to_l2_map()
    assert l3_to_l2 == meta["l3_to_l2"] and l2_names == meta["l2_names"]
    assert list(gen_prompts(False, False)) == meta["gen_prompts_flat"]
    assert list(gen_prompts(True, False)) == meta["gen_prompts_hier"]


def test_gen_prompts_all_modes_match_reference(capsys):
    """data/templates.py:236-297 incl. the descriptive-attribute prompts (a T1 input the reference supports), against
    outputs of the reference's own gen_prompts (tools/export_prompt_data.py -> tests/golden/reference_prompts.json):
    prompts, templates_per_class and the printed per-class preview."""
    import json
    from pathlib import Path
    gold = json.loads((Path(__file__).resolve().parent / "golden" / "reference_prompts.json").read_text())
    for h in (True, False):
        for d in (True, False):
            prompts, tpc = gen_prompts(use_hierarchy=h, use_descriptive=d)
            g = gold[f"hier={h},desc={d}"]
            assert prompts == g["prompts"] and tpc == g["templates_per_class"]
            assert capsys.readouterr().out == g["stdout"]
    assert gen_prompts() == (gold["hier=True,desc=True"]["prompts"], 1)   # the reference's defaults
    capsys.readouterr()


def _have_vocab():
    from aihab_clip_b200.clip.simple_tokenizer import find_vocab
    try:
        find_vocab()
        return True
    except FileNotFoundError:
        return False


@pytest.mark.skipif(not _have_vocab(), reason="CLIP BPE merge table not available on this host")
def test_tokenizer_matches_reference(gold, meta):
    import aihab_clip_b200.clip as clip
    np.testing.assert_array_equal(clip.tokenize(meta["tok_strings"]).numpy(), gold["tok_tokens"])
    np.testing.assert_array_equal(clip.tokenize([meta["tok_long"]], truncate=True).numpy(), gold["tok_truncated"])
    with pytest.raises(RuntimeError):
        clip.tokenize([meta["tok_long"]])
    assert clip.tokenize("a habitat photo of bog.").shape == (1, 77)


def test_load_protocol_and_text_tower_cpu(tmp_path, gold):
    import aihab_clip_b200.clip as clip
    path = tmp_path / "tiny.pt"
    sd = make_state_dict("ViT-tiny/16", 0)
    torch.save(sd, path)
    state, model, preprocess = clip.load(str(path), device="cpu")
    assert set(state.keys()) == set(sd.keys())
    assert model.dtype == torch.float32 and not model.training
    assert model.visual.input_resolution == 64 and tuple(model.visual.proj.shape) == (128, 64)
    assert next(model.parameters()).device.type == "cpu"
    for k in ("visual.proj", "visual.conv1.weight", "visual.transformer.resblocks.1.attn.in_proj_weight"):
        assert torch.equal(state[k], sd[k]), k      # fp16-representable weights survive convert_weights
    xb, x = model.encode_text(torch.from_numpy(gold["tiny16_tok"]))
    np.testing.assert_allclose(xb.detach().numpy(), gold["tiny16_text_before"], atol=2e-4, rtol=0)
    np.testing.assert_allclose(x.detach().numpy(), gold["tiny16_text_emb"], atol=2e-4, rtol=0)
    from aihab_clip_b200.text_head import text_weights_from_tokens
    w = text_weights_from_tokens(model, torch.from_numpy(gold["tok_tokens"][:20]))
    np.testing.assert_allclose(w.numpy(), gold["tiny16_text_w"], atol=2e-5, rtol=0)
    # the PIL preprocess callable is the reference's torchvision pipeline
    from PIL import Image
    from oracle import clip_oracle as O
    img = np.random.default_rng(0).integers(0, 256, (100, 80, 3), dtype=np.uint8)
    np.testing.assert_array_equal(preprocess(Image.fromarray(img)).numpy(), O.clip_preprocess(img, 64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.encode_image(torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError):
        clip.load("ViT-B/16", download_root=str(tmp_path))
    with pytest.raises(RuntimeError):
        clip.load("no-such-model")
    with pytest.raises(NotImplementedError):
        clip.model.build_model({"visual.layer1.0.conv1.weight": torch.zeros(1)})


@pytest.mark.skipif(not _have_vocab(), reason="CLIP BPE merge table not available on this host")
def test_clip_classifier_matches_reference(tmp_path, gold):
    import aihab_clip_b200.clip as clip
    from aihab_clip_b200.text_head import clip_classifier
    path = tmp_path / "tiny.pt"
    torch.save(make_state_dict("ViT-tiny/16", 0), path)
    _, model, _ = clip.load(str(path), device="cpu")
    texts, before, weights = clip_classifier(CS_CLASSNAMES, CS_TEMPLATES, model)
    np.testing.assert_array_equal(texts.numpy(), gold["tiny16_texts"])
    np.testing.assert_allclose(weights.numpy(), gold["tiny16_text_w"], atol=2e-5, rtol=0)
    assert tuple(before.shape) == (1, 20, 64)


def test_cache_dir_naming_matches_reference(meta):
    for cfg, fdir, edir in zip(meta["cache_cfgs"], meta["cache_feature_dirs"], meta["cache_embedding_dirs"]):
        assert str(FC._feature_cache_dir(cfg)) == fdir
        assert str(FC._embedding_cache_dir(cfg, "Test")) == edir
    for name, want in meta["cache_backbone_names"].items():
        assert FC._canonical_backbone_name(name) == want
    assert not FC._feature_cache_exists(Path("/nonexistent"), 1)


def test_evaluation_mirrors_match_reference(gold, meta):
    logits, labels = torch.from_numpy(gold["ev_logits"]), torch.from_numpy(gold["ev_labels"])
    l3_to_l2, n2 = meta["l3_to_l2"], len(meta["l2_names"])
    for red in ("sum", "mean", "logsumexp"):
        got = EV.aggregate_logits_to_l2(logits, l3_to_l2, n2, red)
        np.testing.assert_array_equal(got.numpy(), gold[f"ev_l2_{red}"])   # same op order -> bit-identical on CPU
    np.testing.assert_array_equal(EV.map_l3_targets_to_l2(labels, l3_to_l2).numpy(), gold["ev_l2_targets"])
    c3, i3, p3 = EV.ClassificationTracker().top3_metrics(logits, labels)
    assert int(c3) == int(gold["ev_top3_correct"])
    np.testing.assert_array_equal(i3.numpy(), gold["ev_top3_idx"])
    np.testing.assert_array_equal(p3.numpy(), gold["ev_top3_probs"])
    assert EV.cls_acc(logits, labels, 1) == pytest.approx(float(gold["ev_acc1"]))
    assert EV.cls_acc(logits, labels, 3) == pytest.approx(float(gold["ev_acc3"]))
    with pytest.raises(ValueError):
        EV.aggregate_logits_to_l2(logits[:, :3], l3_to_l2, n2)
    with pytest.raises(ValueError):
        EV.aggregate_logits_to_l2(logits, l3_to_l2, n2, "max")
    with pytest.raises(ValueError):
        EV.L2MetricsAccumulator(l3_to_l2, n2, mode="bogus")


def test_l2_metrics_accumulator(gold, meta):
    from sklearn.metrics import f1_score, matthews_corrcoef
    logits, labels = torch.from_numpy(gold["ev_logits"]), torch.from_numpy(gold["ev_labels"])
    l3_to_l2, n2 = meta["l3_to_l2"], len(meta["l2_names"])
    for mode in ("argmax", "logits"):
        acc = EV.L2MetricsAccumulator(l3_to_l2, n2, reduce="mean", topk=(1, 3), return_confusion_matrix=True, mode=mode)
        acc.update(logits[:40], labels[:40])
        acc.update(logits[40:], labels[40:])
        acc.update(logits[:0], labels[:0])
        m = acc.compute()
        y_true = np.asarray(l3_to_l2)[gold["ev_labels"]]
        if mode == "argmax":
            y_pred = np.asarray(l3_to_l2)[gold["ev_logits"].argmax(1)]
            assert set(m) == {"top1", "f1", "mcc", "cm"}
        else:
            y_pred = gold["ev_l2_mean"].argmax(1)
            assert m["top3"] >= m["top1"]
        assert m["top1"] == pytest.approx((y_true == y_pred).mean())
        assert m["f1"] == pytest.approx(f1_score(y_true, y_pred, average="weighted"))
        assert m["mcc"] == pytest.approx(matthews_corrcoef(y_true, y_pred))
        assert m["cm"].sum() == 64 and m["cm"].shape == (n2, n2)
    empty = EV.L2MetricsAccumulator(l3_to_l2, n2).compute()
    assert empty["top1"] == 0 and empty["f1"] == 0.0 and empty["cm"] is None


@pytest.mark.parametrize("n,world", [(100000, 8), (10, 4), (7, 8), (0, 2), (128, 1)])
def test_shard_ranges_partition_the_index_space(n, world):
    ranges = [shard_range(n, r, world) for r in range(world)]
    covered = [i for lo, hi, _ in ranges for i in range(lo, hi)] if n <= 1000 else None
    if covered is not None:
        assert covered == list(range(n))
    assert ranges[0][0] == 0 and max(hi for _, hi, _ in ranges) == n
    assert all(hi - lo <= per for lo, hi, per in ranges) and len({per for _, _, per in ranges}) == 1


def test_weights_generator_geometry():
    for name, g in GEOMETRIES.items():
        assert g.vision_width % 128 == 0 and g.vision_heads * 64 == g.vision_width
        assert g.tokens == (g.image_resolution // g.vision_patch_size) ** 2 + 1
    sd = make_state_dict("ViT-tiny/14", 1, with_text=False)
    assert sd["visual.conv1.weight"].shape == (256, 3, 14, 14)
    assert all(k.startswith("visual.") for k in sd)
    assert torch.equal(sd["visual.proj"], sd["visual.proj"].half().float())


def test_preferred_batch_fills_whole_waves():
    """aihab_preferred_batch is host arithmetic (no kernel launch): the chosen batch keeps every transformer GEMM
    within a few percent of whole waves of 256-row tile pairs on 148 SMs, and never exceeds the bound."""
    from aihab_clip_b200 import _lib
    lib = _lib.load()
    for tokens, width, bound in [(50, 768, 1024), (197, 768, 256), (257, 1024, 128), (577, 1024, 64), (197, 768, 7)]:
        b = lib.aihab_preferred_batch(tokens, width, bound, 0)
        assert max(1, bound // 2) <= b <= bound
    b = lib.aihab_preferred_batch(50, 768, 1024, 0)
    pair_tiles = -(-b * 50 // 256)
    assert pair_tiles % 74 == 0  # ViT-B/32: 757 images = 148 tile pairs = 2 per SM pair


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): exits 0 without a GPU, prints exactly one
    JSON line with the base contract's keys plus impl / cpu_baseline / e2e, and times the UNMODIFIED reference from
    oracle/_ref when that archive is built (else the torch-operator port)."""
    import json
    import subprocess
    import sys
    from pathlib import Path
    repo = Path(__file__).resolve().parent.parent
    from oracle import build_ref
    build_ref.build(verbose=False)   # no-op where /root/reference is absent
    out = subprocess.run([sys.executable, str(repo / "bench.py"), "--impl", "reference", "--arch", "ViT-tiny/16",
                          "--steps", "2", "--warmup", "1", "--ref-batch", "4"], capture_output=True, text=True,
                         timeout=300, cwd=repo)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["steps"] == 2
    assert d["cpu_baseline"]["cores"] >= 1
    if build_ref.available():
        assert d["cpu_baseline"]["kind"] == "reference" and "oracle/_ref" in d["cpu_baseline"]["sample"]
    else:
        assert d["cpu_baseline"]["kind"] == "port" and "clip_oracle_torch" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # both arms describe the workload with the same `config` object (the driver's same_config check)
    import argparse
    import bench
    from aihab_clip_b200.weights import GEOMETRIES
    a = argparse.Namespace(arch="ViT-tiny/16", classes=20, batch=256)
    assert d["config"] == bench.make_config(a, GEOMETRIES["ViT-tiny/16"], 1)
    # rank != 0 of a torchrun launch exits 0 without work and without output
    import os
    env = dict(os.environ, RANK="1")
    quiet = subprocess.run([sys.executable, str(repo / "bench.py"), "--impl", "reference", "--arch", "ViT-tiny/16"],
                           capture_output=True, text=True, timeout=300, cwd=repo, env=env)
    assert quiet.returncode == 0 and quiet.stdout.strip() == ""


def test_reference_archive_is_the_unmodified_reference():
    """oracle/_ref/reference_path.zip (built here from /root/reference, git-ignored) holds byte-identical copies."""
    import zipfile
    from pathlib import Path
    from oracle import build_ref
    if not build_ref.REF_SRC.is_dir():
        pytest.skip("/root/reference is not present on this box")
    build_ref.build(verbose=False)
    with zipfile.ZipFile(build_ref.ARCHIVE) as z:
        for rel in build_ref.FILES:
            assert z.read(rel) == (build_ref.REF_SRC / rel).read_bytes(), rel
    ignored = (Path(__file__).resolve().parent.parent / ".gitignore").read_text()
    assert "oracle/_ref/" in ignored


def test_model_object_copies_and_engine_fingerprint():
    """the engine fingerprint excludes visual.proj (the engine never reads it); deepcopy / pickle of a model
    work (the ctypes handle is not carried); parameter replacement (.float()) refreshes the cached parameter list."""
    import copy
    import pickle
    import torch
    from aihab_clip_b200.clip.model import build_model
    from aihab_clip_b200.weights import make_state_dict
    m = build_model(make_state_dict("ViT-tiny/16", 0))
    ps = m.visual._engine_params()
    assert all(p is not m.visual.proj for p in ps) and len(ps) == len(list(m.visual.parameters())) - 1
    m.float()
    assert m.visual._engine_params()[0].dtype == torch.float32          # refreshed after _apply
    m2 = copy.deepcopy(m)
    assert m2.visual._engine is None and torch.equal(m2.visual.conv1.weight, m.visual.conv1.weight)
    m3 = pickle.loads(pickle.dumps(m))
    assert torch.equal(m3.visual.proj, m.visual.proj)
    m.visual.invalidate_engine()
    assert m.visual._engine is None
