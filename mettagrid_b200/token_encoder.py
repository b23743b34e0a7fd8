"""Token-policy front end on the device.

``TokenSummary`` stands where the first half of the reference's ``TokenPolicyNet._encode_tokens`` stands
(python/src/mettagrid/policy/token_encoder.py:89-113): position / feature embeddings, value scaling, masking and the
pooled, count-normalised summary per agent row -- computed by ``k_token_summary`` (csrc/mg_gridobs.cu) straight from
the observation rows in HBM instead of through a ``[rows, T, hidden]`` intermediate.  The ``token_mlp`` and the heads
that follow are ordinary torch modules and stay that.  Inference path (forward only).
"""

from __future__ import annotations

import torch

from .sim import BatchedSimulation


class TokenSummary(torch.nn.Module):
    """Same parameters as the reference module's front end: pos_x_embed, pos_y_embed, feature_embed, _feature_scale."""

    def __init__(self, sim: BatchedSimulation, hidden_size: int = 192):
        super().__init__()
        P = sim.program
        self._sim = sim
        self.hidden_size = int(hidden_size)
        max_id = max(P.feature_ids.values(), default=-1)
        n = max(256, max_id + 1)  # token_encoder.py:44-45
        scale = torch.ones(n, dtype=torch.float32)
        for name, fid in P.feature_ids.items():
            scale[fid] = max(float(P.feature_norms[name]), 1.0)
        self.register_buffer("_feature_scale", scale)
        self.pos_x_embed = torch.nn.Embedding(256, self.hidden_size)
        self.pos_y_embed = torch.nn.Embedding(256, self.hidden_size)
        self.feature_embed = torch.nn.Embedding(n, self.hidden_size)
        self.to(sim.device)

    @torch.no_grad()
    def forward(self, observations: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
        s = self._sim
        src = s.observations if observations is None else observations
        if src.dtype != torch.uint8 or not src.is_cuda or not src.is_contiguous() or src.shape[-2:] != (s.num_tokens, 3):
            raise ValueError("observations must be a contiguous CUDA uint8 tensor [..., num_tokens, 3]")
        rows = src.numel() // (s.num_tokens * 3)
        if out is None:
            out = torch.empty((rows, self.hidden_size), dtype=torch.float32, device=s.device)
        w = [self.pos_x_embed.weight, self.pos_y_embed.weight, self.feature_embed.weight, self._feature_scale]
        if any(t.dtype != torch.float32 or not t.is_contiguous() for t in w):
            raise ValueError("embedding tables must be contiguous float32")
        s._check(s._L.mg_token_summary(s._h, src.data_ptr(), rows, w[0].data_ptr(), w[1].data_ptr(), w[2].data_ptr(), w[3].data_ptr(),
                                       self.hidden_size, int(w[2].shape[0]), out.data_ptr(), s._stream()))  # fmt: skip
        return out
