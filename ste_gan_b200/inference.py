"""Generator-only inference (EMGGenerator.generate, models/generator.py:48-50; called at
ste_gan/train.py:394) as a serving loop: weights are folded once, each (batch, frames) shape
is captured as a CUDA graph, utterances are sharded round-robin over ranks with no collective.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import torch

from . import passes
from .dist import round_robin

Tensor = torch.Tensor


class UtteranceGenerator:
    """max_graphs bounds the per-shape CUDA-graph cache (least recently used shape is dropped): each captured (batch,
    frames) shape pins its activations, so serving arbitrary utterance lengths would otherwise grow without limit.
    (The generator is purely
    convolutional with zero padding, so a padded utterance is NOT identical near its end: lengths are never padded
    here - bound the cache instead)."""

    def __init__(self, net_g, precision: str = "bf16", max_graphs: int = 8, trainer=None):
        self.net_g = net_g
        self.dtype = torch.bfloat16 if precision == "bf16" else torch.float32
        self.device = next(net_g.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("UtteranceGenerator: move the generator to a CUDA device first (no CPU path)")
        self.trainer = trainer            # a GanTrainer that owns net_g: its deferred optimiser step is flushed first
        if trainer is not None:
            trainer.flush()
        self.folds = passes.fold_generator(net_g, self.dtype, want_dgrad=False)   # weights are frozen while serving
        self.max_graphs = max(1, int(max_graphs))
        self._graphs: "OrderedDict[Tuple[int, int], tuple]" = OrderedDict()
        self._uses_mode = bool(getattr(net_g, "use_speaking_mode_embedding", False))

    def refold(self) -> None:
        """Call after the generator's weights changed."""
        if self.trainer is not None:
            self.trainer.flush()          # (pipelined step_graph defers the last generator AdamW)
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
        while len(self._graphs) >= self.max_graphs:      # LRU bound
            self._graphs.popitem(last=False)
        su = torch.zeros(batch, frames, unit_dim, device=self.device)
        sess = torch.zeros(batch, device=self.device, dtype=torch.int64)
        mode = torch.zeros(batch, device=self.device, dtype=torch.int64) if self._uses_mode else None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self.generate(su, sess, mode)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = self.generate(su, sess, mode)
        self._graphs[key] = (g, su, sess, mode, out)

    @torch.no_grad()
    def generate_graph(self, speech_units: Tensor, session_ids: Tensor, speaking_mode_ids: Optional[Tensor] = None) -> Tensor:
        """Replay the captured graph for this shape; inputs may be pinned-host tensors.  The returned
        tensor is the graph's static output buffer (overwritten by the next call of the same shape)."""
        B, T, D = speech_units.shape
        if (B, T) not in self._graphs:
            self.capture(B, T, D)
        self._graphs.move_to_end((B, T))
        g, su, sess, mode, out = self._graphs[(B, T)]
        su.copy_(speech_units, non_blocking=True)
        sess.copy_(session_ids, non_blocking=True)
        if mode is not None:
            if speaking_mode_ids is None:
                raise ValueError("generate_graph: this generator uses speaking-mode embeddings - pass speaking_mode_ids")
            mode.copy_(speaking_mode_ids, non_blocking=True)
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
