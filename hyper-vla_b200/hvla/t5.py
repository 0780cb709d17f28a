"""T5-base token embedder on the GPU.

Mirrors what the reference does upstream of ``create_tasks``: the tokenised instruction goes through
``FlaxT5EncoderModel('t5-base')`` and ``last_hidden_state`` becomes ``instruction_dict["language_instruction"]
["token_embedding"]`` (octo/model/components/tokenizers.py:186-211, data/utils/language_tokenizer.py:9-28,
data/simpler/evaluate.py:240-262).  The embeddings stay on the device and can be handed to ``HyperVLA.create_tasks``.

Weights: a HF torch-style state dict (``shared.weight``, ``encoder.block.N.layer.0.SelfAttention.q.weight`` ...; numpy or
torch tensors) or the Flax parameter tree the reference holds (``shared/embedding``, ``.../q/kernel`` [in,out], ...).
"""
from __future__ import annotations

import numpy as np

from . import _native as N

D, H, FF, LAYERS, VOCAB, BUCKETS, MAXDIST = 768, 12, 3072, 12, 32128, 32, 128


def flax_tree_to_state_dict(tree: dict) -> dict:
    """FlaxT5EncoderModel params -> HF torch names ([in,out] kernels transposed to [out,in])."""
    sd = {"shared.weight": np.asarray(tree["shared"]["embedding"])}
    enc = tree["encoder"]
    for l in range(LAYERS):
        blk = enc["block"][str(l)]["layer"]
        att, p = blk["0"]["SelfAttention"], f"encoder.block.{l}.layer."
        for n in ("q", "k", "v", "o"):
            sd[p + f"0.SelfAttention.{n}.weight"] = np.asarray(att[n]["kernel"]).T
        if "relative_attention_bias" in att:
            sd[p + "0.SelfAttention.relative_attention_bias.weight"] = np.asarray(att["relative_attention_bias"]["embedding"])
        sd[p + "0.layer_norm.weight"] = np.asarray(blk["0"]["layer_norm"]["weight"])
        sd[p + "1.DenseReluDense.wi.weight"] = np.asarray(blk["1"]["DenseReluDense"]["wi"]["kernel"]).T
        sd[p + "1.DenseReluDense.wo.weight"] = np.asarray(blk["1"]["DenseReluDense"]["wo"]["kernel"]).T
        sd[p + "1.layer_norm.weight"] = np.asarray(blk["1"]["layer_norm"]["weight"])
    sd["encoder.final_layer_norm.weight"] = np.asarray(enc["final_layer_norm"]["weight"])
    return sd


def relative_bucket(rel: np.ndarray) -> np.ndarray:
    """T5's bidirectional relative-position bucket (32 buckets, max distance 128); rel = key - query."""
    nb = BUCKETS // 2
    out = (rel > 0).astype(np.int64) * nb
    n = np.abs(rel)
    max_exact = nb // 2
    large = max_exact + (np.log(np.maximum(n, 1).astype(np.float32) / max_exact) / np.log(MAXDIST / max_exact) * (nb - max_exact)).astype(np.int64)
    return out + np.where(n < max_exact, n, np.minimum(large, nb - 1))


def pack_t5(sd: dict) -> np.ndarray:
    """One fp32 blob in the layout of include/hvla.h (hvla_t5_encode)."""
    g = lambda k: np.asarray(sd[k].detach().cpu().numpy() if hasattr(sd[k], "detach") else sd[k], np.float32)
    parts = [g("shared.weight").ravel()]
    for l in range(LAYERS):
        p = f"encoder.block.{l}.layer."
        parts += [g(p + "0.layer_norm.weight"), g(p + "0.SelfAttention.q.weight").ravel(), g(p + "0.SelfAttention.k.weight").ravel(),
                  g(p + "0.SelfAttention.v.weight").ravel(), g(p + "0.SelfAttention.o.weight").ravel(), g(p + "1.layer_norm.weight"),
                  g(p + "1.DenseReluDense.wi.weight").ravel(), g(p + "1.DenseReluDense.wo.weight").ravel()]
    parts.append(g("encoder.final_layer_norm.weight"))
    blob = np.concatenate(parts)
    if g("shared.weight").shape != (VOCAB, D):
        raise ValueError("only the t5-base encoder (vocab 32128, d_model 768) is supported")
    return blob


def split_matrices(blob):
    """The GEMM weights of the packed fp32 blob (torch tensor, any device) as split bf16 operands: every [N,K] matrix becomes
    [N,3K] with rows [hi | hi | lo], hi = bf16(w), lo = bf16(w - hi); per layer wqkv | wo | wi | wo2 (include/hvla.h,
    hvla_t5_encode_tc: the activations are laid out [hi | lo | hi], so one GEMM adds hi.hi + lo.hi + hi.lo)."""
    import torch
    layer = D + 3 * D * D + D * D + D + FF * D + D * FF
    out = []
    for l in range(LAYERS):
        base = VOCAB * D + l * layer
        o = base + D
        mats = [(blob[o: o + 3 * D * D], D), (blob[o + 3 * D * D: o + 4 * D * D], D)]
        o += 4 * D * D + D
        mats += [(blob[o: o + FF * D], D), (blob[o + FF * D: o + 2 * FF * D], FF)]
        for w, k in mats:
            w = w.reshape(-1, k)
            hi = w.to(torch.bfloat16)
            lo = (w - hi.float()).to(torch.bfloat16)
            out.append(torch.cat([hi, hi, lo], dim=1).reshape(-1))
    return torch.cat(out)


class T5TokenEmbedder:
    """precision: "bf16x3" (default; tcgen05 GEMMs on split operands, 3e-5 of the fp64 result) or "fp32" (CUDA-core GEMMs in
    the reference's operation order, 4e-6; several times slower)."""

    def __init__(self, weights: dict, device=None, precision: str = "bf16x3"):
        import torch
        if precision not in ("bf16x3", "fp32"):
            raise ValueError("precision must be 'bf16x3' or 'fp32'")
        self.precision = precision
        if not torch.cuda.is_available():
            raise N.HvlaError("no CUDA device: the hvla T5 embedder is CUDA-only")
        self.lib = N.lib()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        sd = weights if "shared.weight" in weights else flax_tree_to_state_dict(weights)
        blob = pack_t5(sd)
        if blob.size != int(self.lib.hvla_t5_blob_elems()):
            raise ValueError("T5 weights do not have the t5-base encoder shapes")
        self.blob = torch.from_numpy(blob).to(self.device)
        self.mat = None
        if precision != "fp32":
            self.mat = split_matrices(self.blob).contiguous()
            if self.mat.numel() != int(self.lib.hvla_t5_mat_elems()):
                raise ValueError("T5 split-weight blob does not match hvla_t5_mat_elems()")
        rb = sd["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"]
        self._rel = np.asarray(rb.detach().cpu().numpy() if hasattr(rb, "detach") else rb, np.float32)      # [32 buckets, 12 heads]
        self._bias, self._ws = {}, None
        # tensor-core path: ~175 dependent launches per call; captured once per (T,S) in a CUDA graph and replayed (as the act step is)
        self.use_graphs = True
        self._graphs = {}

    def _pos_bias(self, S: int):
        import torch
        if S not in self._bias:
            pos = np.arange(S)
            bucket = relative_bucket(pos[None, :] - pos[:, None])                    # [query, key]
            self._bias[S] = torch.from_numpy(np.ascontiguousarray(self._rel[bucket].transpose(2, 0, 1))).to(self.device)
        return self._bias[S]

    def __call__(self, input_ids, attention_mask):
        """(T,S) token ids and mask (numpy or tensors) -> token_embedding (T,S,768) float32 CUDA tensor (asynchronous)."""
        import torch
        ids = torch.as_tensor(np.asarray(input_ids) if not torch.is_tensor(input_ids) else input_ids).to(self.device, torch.int32).contiguous()
        am = torch.as_tensor(np.asarray(attention_mask) if not torch.is_tensor(attention_mask) else attention_mask).to(self.device, torch.int32).contiguous()
        if ids.dim() != 2 or ids.shape != am.shape:
            raise ValueError("input_ids and attention_mask must both be (T,S)")
        T, S = int(ids.shape[0]), int(ids.shape[1])
        if S < 1 or S > 32:
            raise ValueError("the GPU T5 embedder supports 1..32 tokens per instruction (the reference pads to 32)")
        out = torch.empty((T, S, D), dtype=torch.float32, device=self.device)
        if T == 0:
            return out
        if self.precision != "fp32" and self.use_graphs:
            st = self._graphs.get((T, S))
            if st is None:
                st = self._capture(T, S)
            st["ids"].copy_(ids, non_blocking=True)
            st["am"].copy_(am, non_blocking=True)
            st["graph"].replay()
            return st["out"].clone()
        need = self._ws_bytes(T, S)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty((need,), dtype=torch.uint8, device=self.device)
        self._encode(ids, am, T, S, out, self._ws)
        return out

    def _ws_bytes(self, T: int, S: int) -> int:
        return int(self.lib.hvla_t5_tc_workspace_bytes(T, S) if self.precision != "fp32" else self.lib.hvla_t5_workspace_bytes(T, S))

    def _encode(self, ids, am, T, S, out, ws):
        import torch
        stream = int(torch.cuda.current_stream(self.device).cuda_stream)
        pb = self._pos_bias(S)
        with torch.cuda.device(self.device):       # launches go to the current device's context
            if self.precision != "fp32":
                st = self.lib.hvla_t5_encode_tc(stream, self.blob.data_ptr(), self.mat.data_ptr(), pb.data_ptr(), ids.data_ptr(),
                                                am.data_ptr(), T, S, out.data_ptr(), ws.data_ptr(), ws.numel())
                N.check(st, "hvla_t5_encode_tc")
            else:
                st = self.lib.hvla_t5_encode(stream, self.blob.data_ptr(), pb.data_ptr(), ids.data_ptr(), am.data_ptr(), T, S,
                                             out.data_ptr(), ws.data_ptr(), ws.numel())
                N.check(st, "hvla_t5_encode")

    def _capture(self, T: int, S: int) -> dict:
        """Static buffers + captured graph of hvla_t5_encode_tc for this (T,S); at most 4 shapes are kept."""
        import torch
        if len(self._graphs) >= 4:
            self._graphs.pop(next(iter(self._graphs)))
        dev = self.device
        st = {"ids": torch.zeros((T, S), dtype=torch.int32, device=dev), "am": torch.ones((T, S), dtype=torch.int32, device=dev),
              "out": torch.empty((T, S, D), dtype=torch.float32, device=dev),
              "ws": torch.empty((self._ws_bytes(T, S),), dtype=torch.uint8, device=dev)}
        self._pos_bias(S)
        self._encode(st["ids"], st["am"], T, S, st["out"], st["ws"])       # one-time kernel setup outside capture
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.device(dev), torch.cuda.graph(graph):
            self._encode(st["ids"], st["am"], T, S, st["out"], st["ws"])
        st["graph"] = graph
        self._graphs[(T, S)] = st
        return st

def token_to_embedding(model: "T5TokenEmbedder", params, tokens: dict, as_numpy: bool = False):
    """Same name and argument meaning as the reference helper (data/utils/language_tokenizer.py:25-29, called at
    data/simpler/evaluate.py:249-252): ``tokens`` is the tokenizer output ``{"input_ids", "attention_mask"}`` of one or more
    instructions, the result their T5 embeddings (T,S,768).  ``model`` is a :class:`T5TokenEmbedder` (it owns the packed
    weights, so ``params`` is accepted for signature compatibility and ignored).  The reference returns a host array
    (``as_numpy=True``); by default the embeddings stay on the device so that
    ``instruction_dict["language_instruction"]["token_embedding"] = token_to_embedding(...)`` feeds ``create_tasks`` directly."""
    if not isinstance(tokens, dict) or "input_ids" not in tokens or "attention_mask" not in tokens:
        raise ValueError("tokens must be the tokenizer output dict with 'input_ids' and 'attention_mask'")
    ids, am = np.asarray(tokens["input_ids"]), np.asarray(tokens["attention_mask"])
    if ids.ndim == 1:
        ids, am = ids[None], am[None]
    out = model(ids, am)
    return out.cpu().numpy() if as_numpy else out
