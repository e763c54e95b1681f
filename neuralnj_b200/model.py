"""`PhyloATTN` — host-side mirror of the reference model interface (model.py:11-209).

Same constructor (`cfgs.model.*`), the same `state_dict` keys (172 tensors for 6 layers) / shapes / creation order
(so `torch.manual_seed(s); PhyloATTN(cfgs)` equals the reference's default init and shipped
checkpoints load unchanged), and the same inference methods and side effects:

    encode_zxr(batch_input, batch_seq_mask)                      model.py:67-88
    decode_zxr(state, batch_seq_mask, (prev_ij, idx, logits))    model.py:158-209
    aggregate(x_i, x_j, (ii, jj), batchwise_ij_indices=True)     model.py:102-155 (environment.py:829 form)

All arithmetic happens in libnnj's CUDA kernels through the C ABI (include/nnj.h).  The
`nn.Module` here only owns the parameters.  Extra, non-reference entry points:
`rollout_fused` (device-resident encode + R-1 NJ steps) and `merge_state`.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from ._lib import NnjError, check, nnj_config

# fp32: CUDA-core arithmetic everywhere.  bf16x3: dense contractions on tcgen05 as split-bf16 (3 products, ~16 operand mantissa bits):
# the mode that keeps Argmax topologies identical to the fp32 reference and the default.  bf16: the "bf16 encoder" of the north star -
# the encoder's contractions with plain bf16 operands (one product), NJ loop as in bf16x3; scores within 1e-2 relative of the reference,
# topologies NOT guaranteed identical (the headline stays the mode that reaches RF = 0).
PRECISIONS = {"fp32": 0, "bf16x3": 1, "bf16": 2}
DEFAULT_PRECISION = "bf16x3"


class _Attn(nn.Module):
    """Parameter container shaped like Row/ColumnSelfAttention (axial_attention.py:24-28, 161-165)."""

    def __init__(self, d: int):
        super().__init__()
        self.k_proj = nn.Linear(d, d)
        self.v_proj = nn.Linear(d, d)
        self.q_proj = nn.Linear(d, d)
        self.out_proj = nn.Linear(d, d)


class _Ffn(nn.Module):
    """Parameter container shaped like FeedForwardNetwork (msa_modules.py:142-143)."""

    def __init__(self, d: int, f: int):
        super().__init__()
        self.fc1 = nn.Linear(d, f)
        self.fc2 = nn.Linear(f, d)


class _Residual(nn.Module):
    """`NormalizedResidualBlock` key layout: .layer.* then .layer_norm.* (msa_modules.py:93-107)."""

    def __init__(self, layer: nn.Module, d: int):
        super().__init__()
        self.layer = layer
        self.layer_norm = nn.LayerNorm(d)


class _AxialLayer(nn.Module):
    def __init__(self, d: int, f: int):
        super().__init__()
        row, col, ffn = _Attn(d), _Attn(d), _Ffn(d, f)   # RNG order of msa_modules.py:31-50
        self.row_self_attention = _Residual(row, d)
        self.column_self_attention = _Residual(col, d)
        self.feed_forward_layer = _Residual(ffn, d)


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class PhyloATTN(nn.Module):
    def __init__(self, cfgs, precision: Optional[str] = None):
        super().__init__()
        mc = cfgs.model
        if precision is None:      # cfgs.model.precision (an extension key, absent from the reference YAMLs) or the library default
            precision = mc.get("precision", DEFAULT_PRECISION) if hasattr(mc, "get") else DEFAULT_PRECISION
        self.vocab_size = mc.vocab_size
        self.patch_size = mc.patch_size
        self.patch_num = mc.fixed_length // self.patch_size
        self.embed_dim = mc.embed_dim
        self.encoder_attn_layers = mc.encoder_attn_layers
        self.num_enc_heads = mc.num_enc_heads
        self.num_enc_layers = mc.num_enc_layers
        self.dropout = 0.4                      # identity at inference (model.py:23)
        d = self.embed_dim
        self.seq_emb_layers = nn.ModuleList([_AxialLayer(d, 4 * d) for _ in range(self.num_enc_layers)])
        self.embed = nn.Sequential(nn.Linear(self.vocab_size * self.patch_size, d), nn.GELU(), nn.Linear(d, d))
        self.h_linear_last = nn.Linear(d, d)
        self.g_linear_last = nn.Linear(d, d)
        self.g_attn_q = nn.Linear(d, d)
        self.g_attn_k = nn.Linear(d, d)
        self.s_out = nn.Sequential(nn.Linear(d, d), nn.GELU(), nn.Linear(d, 1))
        self.precision = precision
        self._handle = None
        self._handle_key = None
        self._ws = None
        self.batch_input = None
        self.seq_mask = None

    # ------------------------------------------------------------------ plumbing
    def model_params(self):
        return list(self.parameters())

    def forward(self, *a, **k):
        raise NotImplementedError("PhyloATTN has no forward(); use encode_zxr / decode_zxr / aggregate (as the reference)")

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _release(self):
        if self._handle is not None:
            _lib.lib().nnj_model_destroy(self._handle)
            self._handle = None

    def _device(self) -> torch.device:
        return next(self.parameters()).device

    def handle(self):
        """The libnnj weights handle for the current parameter values (rebuilt when they change)."""
        dev = self._device()
        if dev.type != "cuda":
            raise NnjError("PhyloATTN runs on CUDA only: move the model with .to('cuda') (no CPU fallback exists)")
        key = (dev.index, self.precision) + tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._handle is not None and key == self._handle_key:
            return self._handle
        self._release()
        L = _lib.lib()
        if self.precision not in PRECISIONS:
            raise NnjError(f"unknown precision {self.precision!r}; choose from {sorted(PRECISIONS)}")
        cfg = nnj_config(self.embed_dim, self.num_enc_heads, self.num_enc_layers, self.vocab_size, self.patch_size,
                         PRECISIONS[self.precision])
        host = [v.detach().to("cpu", torch.float32).contiguous() for v in self.state_dict().values()]
        n = len(host)
        ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in host])
        numels = (C.c_int64 * n)(*[t.numel() for t in host])
        h = C.c_void_p()
        dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
        check(L.nnj_model_create(C.byref(h), C.byref(cfg), ptrs, numels, n, dev_index), "nnj_model_create")
        self._handle, self._handle_key = h, key
        return h

    def _workspace(self, what: int, B: int, R: int, L: int) -> torch.Tensor:
        need = _lib.lib().nnj_workspace_bytes(self.handle(), what, B, R, L)
        if need < 0:
            check(int(need), "nnj_workspace_bytes")
        if self._ws is None or self._ws.numel() < need or self._ws.device != self._device():
            self._ws = None
            self._ws = torch.empty(int(need), dtype=torch.uint8, device=self._device())
        return self._ws

    @staticmethod
    def _mask_u8(mask: Optional[torch.Tensor], device) -> Optional[torch.Tensor]:
        if mask is None:
            return None
        return mask.to(device=device, dtype=torch.uint8).contiguous()

    # ------------------------------------------------------------------ reference API
    def encode_zxr(self, batch_input, batch_seq_mask=None):
        dev = self._device()
        data = batch_input.to(device=dev, dtype=torch.int8).contiguous()
        B, R, L, E = data.shape
        if E != self.vocab_size:
            raise NnjError(f"encode_zxr: last dim {E} != vocab_size {self.vocab_size}")
        self.patch_num = math.ceil(L / self.patch_size)
        mask = self._mask_u8(batch_seq_mask, dev)
        out = torch.empty(B, R, self.patch_num, self.embed_dim, dtype=torch.float32, device=dev)
        ws = self._workspace(0, B, R, L)
        check(_lib.lib().nnj_encode(self.handle(), _ptr(data), _ptr(mask), B, R, L, _ptr(out), _ptr(ws), ws.numel(), _stream()),
              "nnj_encode")
        return out

    def decode_zxr(self, batch_input, batch_seq_mask=None, indices_to_prev_info=None):
        dev = self._device()
        state = batch_input.to(device=dev, dtype=torch.float32).contiguous()
        B, Rp, Cc = state.shape[:3]
        actions_ij_prev, score_indices_to_prev, logits_prev = indices_to_prev_info
        mask = self._mask_u8(batch_seq_mask[:, ::self.patch_size] if batch_seq_mask is not None else None, dev)
        self.batch_input = state
        self.seq_mask = None if batch_seq_mask is None else ~batch_seq_mask[:, None, ::self.patch_size]
        L = _lib.lib()
        ws = self._workspace(1, B, Rp, Cc)
        if logits_prev is None:
            scores = torch.empty(B, Rp * (Rp - 1) // 2, dtype=torch.float32, device=dev)
            check(L.nnj_pair_scores_full(self.handle(), _ptr(state), _ptr(mask), B, Rp, Cc, _ptr(scores), _ptr(ws), ws.numel(), _stream()),
                  "nnj_pair_scores_full")
        else:
            prev = actions_ij_prev.to(device=dev, dtype=torch.int32).contiguous()
            lp = logits_prev.to(device=dev, dtype=torch.float32).contiguous()
            if score_indices_to_prev is None:
                scores = torch.empty(B, Rp * (Rp - 1) // 2, dtype=torch.float32, device=dev)
                check(L.nnj_pair_scores_incr(self.handle(), _ptr(state), _ptr(mask), B, Rp, Cc, _ptr(prev), _ptr(lp), _ptr(scores),
                                             _ptr(ws), ws.numel(), _stream()), "nnj_pair_scores_incr")
            else:
                # the reference's own formulation: Rp new pairs (i*, r) sorted, then the caller's gather map
                r = torch.arange(Rp, device=dev, dtype=torch.int32).unsqueeze(0).expand(B, Rp)
                a = prev[:, :1].expand(B, Rp)
                pi, pj = torch.minimum(a, r).contiguous(), torch.maximum(a, r).contiguous()
                new = torch.empty(B, Rp, dtype=torch.float32, device=dev)
                check(L.nnj_pair_scores_list(self.handle(), _ptr(state), _ptr(mask), B, Rp, Cc, _ptr(pi), _ptr(pj), Rp, _ptr(new),
                                             _ptr(ws), ws.numel(), _stream()), "nnj_pair_scores_list")
                scores = torch.gather(torch.cat([lp, new], dim=-1), 1, score_indices_to_prev.to(dev))
        return {"logits": scores, "distance": scores}

    def aggregate(self, x_i, x_j, ij_indices, batchwise_ij_indices=False):
        """Merged-node embedding of one pair per tree against `self.batch_input` (environment.py:822-831)."""
        if not batchwise_ij_indices:
            raise NotImplementedError("aggregate: only the batchwise_ij_indices=True form (PhyInferEnv.step) crosses the "
                                      "boundary; pair scoring goes through decode_zxr")
        if self.batch_input is None:
            raise NnjError("aggregate: call decode_zxr first (the reference reads self.batch_input set there)")
        state = self.batch_input
        B, Rp, Cc = state.shape[:3]
        ii, jj = ij_indices
        ij = torch.stack([ii, jj], dim=1).to(device=state.device, dtype=torch.int32).contiguous()
        out = torch.empty(B, 1, Cc, self.embed_dim, dtype=torch.float32, device=state.device)
        ws = self._workspace(1, B, Rp, Cc)
        check(_lib.lib().nnj_aggregate(self.handle(), _ptr(state), B, Rp, Cc, _ptr(ij), _ptr(out), _ptr(ws), ws.numel(), _stream()),
              "nnj_aggregate")
        return out

    # ------------------------------------------------------------------ fused paths (not in the reference)
    def merge_state(self, state: torch.Tensor, ij: torch.Tensor) -> torch.Tensor:
        """Tensor half of PhyInferEnv.step: slot i <- aggregate(i, j), slot j removed (environment.py:760-835)."""
        state = state.contiguous()
        B, Rp, Cc = state.shape[:3]
        ij = ij.to(device=state.device, dtype=torch.int32).contiguous()
        out = torch.empty(B, Rp - 1, Cc, self.embed_dim, dtype=torch.float32, device=state.device)
        ws = self._workspace(1, B, Rp, Cc)
        check(_lib.lib().nnj_merge(self.handle(), _ptr(state), B, Rp, Cc, _ptr(ij), _ptr(out), _ptr(ws), ws.numel(), _stream()),
              "nnj_merge")
        return out

    def rollout_fused(self, batch_input=None, batch_seq_mask=None, gumbel: Optional[torch.Tensor] = None,
                      state: Optional[torch.Tensor] = None, want_logits: bool = False,
                      forced: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
        """Encode + all R-1 NJ steps on the device (finetune_rl_search.py:107-175 without host round trips).

        Returns (merges int32 [B,R-1,2], selected_logp fp32 [B,R-1], logits trace [B, sum_t P_t] or None).
        `gumbel` fp32 [B,R-1,R(R-1)/2] switches selection to argmax(logits + gumbel) (sampling);
        `state` supplies a pre-computed encoder output shared by several rollouts (Search mode);
        `forced` int32 [B,R-1,2] replays the given actions instead of selecting (teacher forcing, train.py:116-119).
        """
        dev = self._device()
        L = _lib.lib()
        mask = self._mask_u8(batch_seq_mask, dev)
        if state is not None:
            state = state.to(device=dev, dtype=torch.float32).contiguous()
            B, R, Cc = state.shape[:3]
            Ls = Cc
        else:
            data = batch_input.to(device=dev, dtype=torch.int8).contiguous()
            B, R, Ls, _ = data.shape
            self.patch_num = math.ceil(Ls / self.patch_size)
        merges = torch.empty(B, R - 1, 2, dtype=torch.int32, device=dev)
        if forced is not None:
            if gumbel is not None:
                raise NnjError("rollout_fused: `forced` and `gumbel` exclude each other")
            if tuple(forced.shape) != (B, R - 1, 2):
                raise NnjError("rollout_fused: forced must be [B, R-1, 2]")
            merges.copy_(forced)                 # NNJ_SELECT_FORCED reads the actions from the merge-list buffer
        slp = torch.empty(B, R - 1, dtype=torch.float32, device=dev)
        trace = None
        if want_logits:
            tot = sum(n * (n - 1) // 2 for n in range(2, R + 1))
            trace = torch.empty(B, tot, dtype=torch.float32, device=dev)
        mode = 2 if forced is not None else 0
        if gumbel is not None:
            gumbel = gumbel.to(device=dev, dtype=torch.float32).contiguous()
            if tuple(gumbel.shape) != (B, R - 1, R * (R - 1) // 2):
                raise NnjError("rollout_fused: gumbel must be [B, R-1, R(R-1)/2]")
            mode = 1
        ws = self._workspace(2, B, R, Ls)
        if state is not None:
            check(L.nnj_rollout_from_state(self.handle(), _ptr(state), _ptr(mask), B, R, Ls, mode, _ptr(gumbel), _ptr(merges),
                                           _ptr(trace), _ptr(slp), _ptr(ws), ws.numel(), _stream()), "nnj_rollout_from_state")
        else:
            check(L.nnj_rollout(self.handle(), _ptr(data), _ptr(mask), B, R, Ls, mode, _ptr(gumbel), _ptr(merges), _ptr(trace),
                                _ptr(slp), _ptr(ws), ws.numel(), _stream()), "nnj_rollout")
        return merges, slp, trace

    def rollout_host(self, data_host: torch.Tensor, mask_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Host-buffer entry point (nnj_rollout_host): int8 MSA on the host in, merge lists on the host out."""
        data_host = data_host.to(torch.int8).contiguous()
        B, R, Ls, _ = data_host.shape
        mask_u8 = None if mask_host is None else mask_host.to(torch.uint8).contiguous()
        merges = torch.empty(B, R - 1, 2, dtype=torch.int32, pin_memory=True)
        check(_lib.lib().nnj_rollout_host(self.handle(), _ptr(data_host), _ptr(mask_u8), B, R, Ls, 0, C.c_void_p(0), _ptr(merges),
                                          C.c_void_p(0)), "nnj_rollout_host")
        return merges
