"""Generator-only inference (EMGGenerator.generate, models/generator.py:48-50; called at
ste_gan/train.py:394) as a serving loop: weights are folded once, each (batch, frames) shape
is captured as a CUDA graph, utterances are sharded round-robin over ranks with no collective.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import passes
from .dist import round_robin

Tensor = torch.Tensor


class UtteranceGenerator:
    def __init__(self, net_g, precision: str = "bf16"):
        self.net_g = net_g
        self.dtype = torch.bfloat16 if precision == "bf16" else torch.float32
        self.device = next(net_g.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("UtteranceGenerator: move the generator to a CUDA device first (no CPU path)")
        self.folds = passes.fold_generator(net_g, self.dtype, want_dgrad=False)   # weights are frozen while serving
        self._graphs: Dict[Tuple[int, int], tuple] = {}

    def refold(self) -> None:
        """Call after the generator's weights changed."""
        self.folds = passes.fold_generator(self.net_g, self.dtype, want_dgrad=False)
        self._graphs.clear()

    @torch.no_grad()
    def generate(self, speech_units: Tensor, session_ids: Tensor, speaking_mode_ids: Optional[Tensor] = None) -> Tensor:
        """[B,T,D] units -> [B,16T,C] fp32 EMG (eager launches)."""
        x, _ = passes.generator_forward(self.net_g, speech_units, session_ids, speaking_mode_ids, self.dtype, False,
                                        folds=self.folds)
        return x

    @torch.no_grad()
    def capture(self, batch: int, frames: int, unit_dim: int) -> None:
        key = (batch, frames)
        su = torch.zeros(batch, frames, unit_dim, device=self.device)
        sess = torch.zeros(batch, device=self.device, dtype=torch.int64)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self.generate(su, sess)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = self.generate(su, sess)
        self._graphs[key] = (g, su, sess, out)

    @torch.no_grad()
    def generate_graph(self, speech_units: Tensor, session_ids: Tensor) -> Tensor:
        """Replay the captured graph for this shape; inputs may be pinned-host tensors.  The returned
        tensor is the graph's static output buffer (overwritten by the next call of the same shape)."""
        B, T, D = speech_units.shape
        if (B, T) not in self._graphs:
            self.capture(B, T, D)
        g, su, sess, out = self._graphs[(B, T)]
        su.copy_(speech_units, non_blocking=True)
        sess.copy_(session_ids, non_blocking=True)
        g.replay()
        return out

    def serve(self, utterances: List[Tuple[Tensor, Tensor]], rank: int = 0, world: int = 1) -> Dict[int, Tensor]:
        """Generate this rank's share (round-robin) of a list of (units [T,D], session_id) utterances;
        returns {utterance index: EMG [16T, C] on the host}."""
        out = {}
        for i in round_robin(rank, world, len(utterances)):
            su, sid = utterances[i]
            y = self.generate_graph(su.unsqueeze(0), sid.reshape(1))
            out[i] = y[0].to("cpu", non_blocking=False)
        return out
