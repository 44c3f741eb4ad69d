"""Host-to-device clip streaming for inference loops (the `data.cuda(); model(data)` loop of the reference's
evaluate()/eval drivers, training/train.py:646-688, evaluation/evaluate.py:275-288).

`ClipStreamer` overlaps the H2D copy of batch i+1 (a side stream, two device staging buffers) with the forward of
batch i and returns each batch's logits in pinned host memory.  Results are identical to calling
`model(batch.cuda())[0].cpu()` batch by batch; only the copies are pipelined.
"""
from __future__ import annotations

from typing import Iterable, Iterator, List, Optional

import torch

Tensor = torch.Tensor


class ClipStreamer:
    def __init__(self, model, device: Optional[torch.device] = None):
        self.model = model
        self.device = device or next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("ClipStreamer needs the model on a CUDA device (no CPU fallback)")
        self.copy_stream = torch.cuda.Stream(self.device)
        self._stage: List[Optional[Tensor]] = [None, None]
        self._staged = [torch.cuda.Event(), torch.cuda.Event()]      # H2D of slot k finished
        self._consumed = [torch.cuda.Event(), torch.cuda.Event()]    # forward that read slot k finished
        self._out: List[Optional[Tensor]] = [None, None]
        self._done = [torch.cuda.Event(), torch.cuda.Event()]
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _tower_engine(self):
        visual = getattr(self.model, "visual", None)
        return visual.engine(self.device) if visual is not None and hasattr(visual, "engine") else None

    def close(self) -> None:
        """Release the staging slots (and the graphs the tower captured on them).  Called on garbage collection too."""
        eng = None
        for k in (0, 1):
            if self._stage[k] is not None:
                try:
                    eng = eng or self._tower_engine()
                    if eng is not None:
                        eng.unregister_input(self._stage[k])
                except Exception:          # interpreter shutdown / model already gone: nothing left to release
                    pass
                self._stage[k] = None

    def __del__(self):
        self.close()

    def _upload(self, k: int, host: Tensor) -> None:
        if self._stage[k] is None or self._stage[k].shape != host.shape or self._stage[k].dtype != host.dtype:
            if self._stage[k] is not None:                 # a re-shaped slot (ragged tail batch): drop the old registration
                eng = self._tower_engine()
                if eng is not None:
                    eng.unregister_input(self._stage[k])
            self._stage[k] = torch.empty(host.shape, device=self.device, dtype=host.dtype)   # fp32 (B,3,T,H,W) or uint8 (B,T,H,W,3)
            # the caching allocator may hand back a block whose last use on the compute stream is still in flight
            self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
            # the two staging slots live as long as the streamer: the tower replays a graph captured on the slot itself
            # instead of copying it into its own static clip buffer first
            eng = self._tower_engine()
            if eng is not None:
                eng.register_input(self._stage[k])
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self._consumed[k])            # slot free again?
            self._stage[k].copy_(host, non_blocking=True)
            self._staged[k].record(self.copy_stream)
        self.h2d_bytes += host.numel() * host.element_size()

    def _forward(self, k: int) -> None:
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._staged[k])
        logits = self.model(self._stage[k])[0]
        self._consumed[k].record(cur)
        if self._out[k] is None or self._out[k].shape != logits.shape:
            self._out[k] = torch.empty(logits.shape, dtype=logits.dtype).pin_memory()
        self._out[k].copy_(logits, non_blocking=True)
        self._done[k].record(cur)
        self.d2h_bytes += logits.numel() * logits.element_size()

    @torch.no_grad()
    def run(self, host_batches: Iterable[Tensor]) -> Iterator[Tensor]:
        """Yields the logits (pinned host tensor, valid until two more batches were consumed) of every batch."""
        it = iter(host_batches)
        try:
            nxt = next(it)
        except StopIteration:
            return
        self._upload(0, nxt)
        i = 0
        pending = None
        while True:
            k = i & 1
            try:
                nxt = next(it)
                self._upload(k ^ 1, nxt)      # copy of batch i+1 overlaps the forward of batch i
                more = True
            except StopIteration:
                more = False
            self._forward(k)
            if pending is not None:
                self._done[pending].synchronize()
                yield self._out[pending]
            pending = k
            if not more:
                break
            i += 1
        self._done[pending].synchronize()
        yield self._out[pending]
