"""Generate tests/golden/input_stage.pt by running the REFERENCE's own VideoDataset (eval branch) in the build container.

    python oracle/make_golden_input.py

PyAV is not installed, so a fake `av` module serves the synthetic uint8 frames of oracle/input_oracle.make_frames through
the three calls the reference makes (av.open(...).decode(video=0) -> frame.pts / frame.to_rgb().to_ndarray(), dataset.py:
86-92,118).  Everything after decoding is the reference's code: `first_view` = what __getitem__ returns (frames[0]);
`views` = its _generate_spatial_crops / _generate_temporal_crops applied to the same resized tensor.
"""
from __future__ import annotations

import sys
import tempfile
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import fixtures as fx  # noqa: E402
from oracle import input_oracle as io  # noqa: E402

REF = Path("/root/reference")
_STORE = {}


class _Frame:
    def __init__(self, pts, arr):
        self.pts, self._arr = pts, arr

    def to_rgb(self):
        return self

    def to_ndarray(self):
        return self._arr


class _Container:
    def __init__(self, path):
        self.frames = _STORE[Path(path).name]

    def decode(self, video=0):
        # served out of order: the reference sorts by pts (dataset.py:91)
        idx = list(range(self.frames.shape[0]))[::-1]
        return [_Frame(i * 512, self.frames[i].numpy()) for i in idx]

    def close(self):
        pass


def main():
    fake_av = types.ModuleType("av")
    fake_av.open = lambda path: _Container(path)
    sys.modules["av"] = fake_av
    sys.path.insert(0, str(REF))
    from video_dataset.dataset import VideoDataset  # the reference's class, unmodified
    import torch.nn.functional as F
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, (T0, H0, W0, S, nf, rate, nsv, ntv) in io.INPUT_CASES.items():
            frames = io.make_frames(name)
            _STORE[f"{name}.mp4"] = frames
            lst = Path(tmp) / f"{name}.csv"
            lst.write_text(f"{name}.mp4,1\n")
            ds = VideoDataset(list_path=str(lst), data_root=tmp, num_spatial_views=nsv, num_temporal_views=ntv,
                              random_sample=False, num_frames=nf, sampling_rate=rate, spatial_size=S,
                              mean=torch.tensor(io.MEAN), std=torch.tensor(io.STD), is_train=False)
            first, label, vidname = ds[0]
            # all views: the reference's crop methods on the tensor __getitem__ builds at dataset.py:117-133
            x = frames.float() / 255.
            x = ((x - ds.mean) / ds.std).permute(3, 0, 1, 2)
            if x.size(-2) < x.size(-1):
                nw, nh = x.size(-1) * S // x.size(-2), S
            else:
                nh, nw = x.size(-2) * S // x.size(-1), S
            x = F.interpolate(x, size=(nh, nw), mode="bilinear", align_corners=False)
            views = sum([ds._generate_temporal_crops(v) for v in ds._generate_spatial_crops(x)], [])
            assert torch.equal(views[0], first)
            out[name] = {"first_view": first.clone(), "views": torch.stack(views).clone(), "label": label, "vidname": vidname}
            print(name, tuple(first.shape), len(views))
    torch.save(out, fx.GOLDEN / "input_stage.pt")


if __name__ == "__main__":
    main()
