"""CPU: pin the oracle (oracle/vita_oracle.py) against outputs of the REFERENCE ITSELF (tests/golden/*.pt, produced
by oracle/make_golden.py from /root/reference in the build container), and check the drop-in's host-side contract
(state_dict layout, frozen set, tokenisation) against what the reference reported."""
import pytest
import torch

from tests.util import build_model, fx, load_golden, rel_err

from oracle import vita_oracle as vo

FAST = ["tiny", "b16_t8", "kapt_b16_t8", "tiny_global_only", "tiny_no_local", "tiny_no_summary"]
SLOW = ["l14_t16", "b16_t16_k400", "b16_t8_b8"]


@pytest.mark.parametrize("name", FAST + SLOW)
def test_state_dict_contract(name):
    g = load_golden(name)
    model, kw, B = build_model(name)
    sd = model.state_dict()
    assert {k: list(v.shape) for k, v in sd.items()} == g["state_dict_shapes"]
    assert list(sd.keys()) == list(g["state_dict_shapes"].keys())          # same order, too
    assert sorted(n for n, p in model.named_parameters() if p.requires_grad) == g["requires_grad"]
    assert len(model.tokenized_prompts) == len(g["tokenized_prompts"])
    for a, b in zip(model.tokenized_prompts, g["tokenized_prompts"]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("name", FAST + SLOW)
def test_oracle_matches_reference(name):
    g = load_golden(name)
    model, kw, B = build_model(name)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    inp = fx.make_inputs(name)
    kapt = bool(kw.get("text_prompt_init"))
    cntn = model.prompt_learner.context_prompt_learner.cntn_embeds if kapt else None
    taps = {}
    out = vo.vita_forward(sd, inp["video"], model.tokenized_prompts, num_frames=kw["num_frames"],
                          num_heads=kw.get("num_heads", 12), num_layers=kw.get("num_layers", 12),
                          text_heads=kw.get("text_transformer_heads", 8), text_layers=kw.get("text_transformer_layers", 12),
                          n_ctx=8, kapt=kapt, cntn_embeds=cntn, taps=taps,
                          use_summary=kw.get("use_summary_token", False), use_local=kw.get("use_local_prompts", False))
    # same algorithm, same fp32 weights, same CPU: only summation-order noise is allowed
    assert rel_err(out["logits"], g["logits"]) < 1e-4
    assert rel_err(out["video_features"], g["video_features"]) < 1e-4
    if g["summary"] is None:
        assert out["summary"] is None          # use_summary_token=False: the reference returns (cls_x, None)
    else:
        assert rel_err(out["summary"], g["summary"]) < 1e-4
    assert rel_err(out["text_features"], g["text_features"]) < 1e-4
    assert torch.equal(out["logits"].argmax(-1), g["logits"].argmax(-1))
    if g["layer_cls"] is not None:
        assert rel_err(torch.stack(taps["layer_cls"]), g["layer_cls"]) < 1e-4
        assert rel_err(torch.stack(taps["layer_summary"]).flatten(1, 2), g["layer_summary"].flatten(1, 2)) < 1e-4
    if g["logits_mt"] is not None:
        mt = vo.memory_head(sd, out["text_features"], inp["memory"], kw["num_classes"])
        assert rel_err(mt, g["logits_mt"]) < 1e-4


def test_oracle_zeroshot_head_consistent():
    """zero-shot branch == prompt branch when every class has one description (VitaCLIP_model.py:296-306)."""
    g = load_golden("tiny")
    v = g["video_features"]
    ls = torch.tensor(2.65926)
    tf = torch.randn(3, v.shape[1], generator=torch.Generator().manual_seed(0))
    a = vo.zeroshot_head(v, tf, ls)
    b, _ = vo.logit_head(v, [t[None] for t in tf], ls)
    assert torch.allclose(a, b, atol=1e-6)


def test_input_stage_oracle_matches_reference_dataset():
    """oracle/input_oracle.py against the reference's own VideoDataset eval branch (oracle/make_golden_input.py)."""
    import torch
    from oracle import input_oracle as io
    gold = torch.load(fx.GOLDEN / "input_stage.pt", map_location="cpu", weights_only=False)
    for name, (T0, H0, W0, S, nf, rate, nsv, ntv) in io.INPUT_CASES.items():
        views = io.eval_views(io.make_frames(name), S, nf, rate, nsv, ntv)
        ref = gold[name]["views"]
        assert len(views) == ref.shape[0] == nsv * ntv
        for v, r in zip(views, ref):
            assert v.shape == r.shape and torch.equal(v, r), name


def test_input_stage_geometry_matches_oracle():
    """EvalInputStage's integer geometry (resize target, crop offsets, temporal starts) reproduces the oracle's crops."""
    import torch
    from gava_clip_b200.input_stage import EvalInputStage, aggregate_views
    from oracle import input_oracle as io
    for name, (T0, H0, W0, S, nf, rate, nsv, ntv) in io.INPUT_CASES.items():
        st = EvalInputStage(S, nf, rate, nsv, ntv, io.MEAN, io.STD)
        nh, nw = st.resized_hw(H0, W0)
        idx = torch.arange(T0 * nh * nw, dtype=torch.float32).view(1, T0, nh, nw)        # position-coded tensor
        want = sum([io.temporal_crops(v, nf, rate, ntv) for v in io.spatial_crops(idx, S, nsv)], [])
        got = []
        for (h0, w0) in st.spatial_offsets(nh, nw):
            for t0 in st.temporal_starts(T0):
                ts = [min(t0 + i * rate, T0 - 1) for i in range(nf)]
                got.append(idx[:, ts, h0:h0 + S, w0:w0 + S])
        assert len(got) == len(want) and all(torch.equal(a, b) for a, b in zip(got, want)), name
    lg = torch.randn(6, 3)
    assert torch.allclose(aggregate_views(lg), lg.softmax(-1).mean(0))
    assert torch.allclose(aggregate_views(lg, n_views=3), lg.softmax(-1).view(2, 3, 3).mean(1))


def test_oracle_memory_and_nte_heads_match_reference():
    """memory_head / nte_head (VitaCLIP_model.py:311-343, 347-391) against the reference's logits_mt / logits_vm goldens."""
    import torch
    from oracle import vita_oracle as vo
    from tests.util import build_model, load_golden
    g = load_golden("kapt_b16_t8")
    model, kw, _ = build_model("kapt_b16_t8")
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    inp = fx.make_inputs("kapt_b16_t8")
    with torch.no_grad():
        v, summary = vo.vision_encoder(sd, inp["video"], num_frames=8, num_heads=12, num_layers=12)
        assert float((vo.nte_head(sd, summary, inp["video_nte"]) - g["logits_vm"]).abs().max()) < 1e-4
        assert float((vo.memory_head(sd, g["text_features"], inp["memory"], 3) - g["logits_mt"]).abs().max()) < 1e-3
