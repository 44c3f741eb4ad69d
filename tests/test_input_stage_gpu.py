"""GPU: the fused input-stage kernel (gava_preprocess_u8_fwd) against the reference dataset's outputs (golden fixture made
by oracle/make_golden_input.py) and against the CPU oracle on a full-size frame; then through the model."""
import pytest
import torch

from tests.util import build_model, fx, rel_err

pytestmark = pytest.mark.gpu
TOL_ABS = 2e-6      # fp32; differences come from FMA contraction only (same weights, same operation order)


def test_views_match_reference_dataset(cuda):
    from gava_clip_b200.input_stage import EvalInputStage
    from oracle import input_oracle as io
    gold = torch.load(fx.GOLDEN / "input_stage.pt", map_location="cpu", weights_only=False)
    for name, (T0, H0, W0, S, nf, rate, nsv, ntv) in io.INPUT_CASES.items():
        st = EvalInputStage(S, nf, rate, nsv, ntv, io.MEAN, io.STD)
        frames = io.make_frames(name).to(cuda)
        views = st.views(frames).cpu()
        ref = gold[name]["views"]
        assert views.shape == ref.shape, name
        assert float((views - ref).abs().max()) < TOL_ABS, name
        assert float((st(frames).cpu() - ref[0]).abs().max()) < TOL_ABS, name


def test_full_size_video_and_model_call(cuda):
    """320x240 decoder frames -> 224^2 clip (oracle), and the clip drives the model like a loader batch."""
    from gava_clip_b200.input_stage import EvalInputStage
    from oracle import input_oracle as io
    g = torch.Generator().manual_seed(3)
    frames = torch.randint(0, 256, (20, 240, 320, 3), dtype=torch.uint8, generator=g)
    st = EvalInputStage(224, 8, 2, 1, 1, io.MEAN, io.STD)
    ref = io.eval_views(frames, 224, 8, 2, 1, 1)[0]
    got = st(frames.to(cuda)).cpu()
    assert float((got - ref).abs().max()) < TOL_ABS
    model, kw, B = build_model("b16_t8", device=cuda)
    batch = st.batch([frames.to(cuda), frames.flip(0).contiguous().to(cuda)])
    with torch.no_grad():
        a = model(batch)[0]
        b = model(torch.stack([ref, io.eval_views(frames.flip(0).contiguous(), 224, 8, 2, 1, 1)[0]]).to(cuda))[0]
    assert rel_err(a, b) < 1e-3

    with pytest.raises(RuntimeError):
        st(frames)                     # CPU tensor: no fallback
