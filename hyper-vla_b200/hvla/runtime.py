"""Device-side state of one HyperVLA model on one B200: packed parameter blobs, workspace,
pinned staging buffers, and the calls into libhvla.so.  PyTorch is used only for device
memory, streams and pinned host memory (plumbing); all compute is in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _native as N
from . import config as Cfg
from . import metadata as M
from . import params as P


def _torch():
    import torch
    return torch


class Runtime:
    def __init__(self, params: dict, precision: str = "bf16", device: Optional[object] = None, spec: "M.HeadSpec" = M.MIX):
        torch = _torch()
        self.spec = spec                # action head variant: README mix head, or DiscreteActionHead on 4 / 28 readout tokens
        self.ngp = M.n_generated_padded(spec)
        if precision not in ("bf16", "fp32", "fp32x3"):
            raise ValueError("precision must be 'bf16', 'fp32' or 'fp32x3' (fp32-class accuracy on the tensor cores, split bf16 operands)")
        if not torch.cuda.is_available():
            raise N.HvlaError("no CUDA device: the hvla hot path is CUDA-only (there is no CPU fallback)")
        self.lib = N.lib()
        self.precision = precision
        self.dtype = {"bf16": N.HVLA_BF16, "fp32": N.HVLA_F32, "fp32x3": N.HVLA_BF16X3}[precision]
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.tdtype = torch.bfloat16 if precision == "bf16" else torch.float32
        self._ws = None
        self._ws_key = (0, 0)
        self._pinned = {}
        self._graphs = {}
        self.use_graphs = True          # replay a captured CUDA graph of the act step instead of ~110 eager launches
        self._profiling = False
        self.replayed_launches = 0      # kernels launched through graph replays (the C counter only sees eager launches)
        self.graph_captures = 0         # CUDA-graph captures so far (a task switch through generate_rows must not add one)
        self.upload(params)

    def _on_device(self):
        """Make this runtime's device current around library calls (launches go to the current device's context)."""
        return _torch().cuda.device(self.device)

    # ---- parameters -------------------------------------------------------------------------------
    def upload(self, params: dict) -> None:
        torch = _torch()
        dev = self.device
        hn = P.pack_hn_blob(params)
        assert hn.size == self.lib.hvla_hn_blob_elems()
        self.hn_blob = torch.from_numpy(hn).to(dev)
        self.hn_blob_f16 = self.hn_blob.to(torch.float16) if self.precision == "bf16" else None
        W, b = P.pack_heads(params, self.spec)
        self.heads_w = torch.from_numpy(W).to(self.tdtype).to(dev)
        self.heads_b = torch.from_numpy(b).to(dev)
        del W
        self.dino_vec, self.dino_mat = self._dino_blobs(*P.pack_dino(params, transposed=(self.precision == "bf16")))
        self._params_id = id(params)
        self._graphs.clear()            # captured graphs hold pointers into the previous blobs

    def _dino_blobs(self, vec, mat):
        """(vec, matrix) host blobs of P.pack_dino* -> device tensors in this runtime's matrix format."""
        torch = _torch()
        assert vec.size == self.lib.hvla_dino_vec_elems() and mat.size == self.lib.hvla_dino_mat_elems()
        if self.precision == "fp32x3":          # [hi | hi | lo] bf16 planes of every transposed matrix
            m = torch.from_numpy(P.split_matrices_x3(mat).view(np.int16)).view(torch.bfloat16).to(self.device)
        else:
            m = torch.from_numpy(mat).to(self.tdtype).to(self.device)
        return torch.from_numpy(vec).to(self.device), m

    # ---- scratch ------------------------------------------------------------------------------------
    def workspace(self, B: int, T: int):
        torch = _torch()
        kb, kt = self._ws_key
        if self._ws is None or B > kb or T > kt:
            nb, nt = max(B, kb), max(T, kt)
            nbytes = int(self.lib.hvla_workspace_bytes(nb, nt, self.dtype))
            self._ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            self._ws_key = (nb, nt)
        ptr = (self._ws.data_ptr() + 255) & ~255
        return ptr, self._ws.numel() - (ptr - self._ws.data_ptr())

    def pinned(self, name: str, shape, dtype):
        torch = _torch()
        t = self._pinned.get(name)
        n = int(np.prod(shape))
        if t is None or t.numel() < n or t.dtype != dtype:
            t = torch.empty(n, dtype=dtype).pin_memory()
            self._pinned[name] = t
        return t[:n].view(*shape)

    def launch_count(self) -> int:
        """Kernels of libhvla launched so far on behalf of this runtime (eager + replayed through CUDA graphs)."""
        return int(self.lib.hvla_launch_count()) + self.replayed_launches

    def stream(self) -> int:
        return int(_torch().cuda.current_stream(self.device).cuda_stream)

    # ---- generate -------------------------------------------------------------------------------------
    def generate(self, token_embedding, attention_mask, init_cls, lang_pad=None, rows=None, into=None):
        """-> (weights [T, NGP] device tensor, ctx_emb [T,128] device tensor).
        ``rows`` + ``into=(weights [T_max,NGP], ctx [T_max,128])``: regenerate only those rows of the persistent buffers in
        place (hvla_generate_rows): the buffers keep their addresses, so CUDA graphs captured over them stay valid."""
        torch = _torch()
        dev = self.device
        tok = torch.as_tensor(np.ascontiguousarray(token_embedding, dtype=np.float32) if not torch.is_tensor(token_embedding)
                              else token_embedding).to(dev, torch.float32).contiguous()
        T = int(tok.shape[0])
        if tuple(tok.shape[1:]) != (Cfg.LANG_TOKENS, Cfg.LANG_DIM):
            raise ValueError(f"token_embedding must be (T,{Cfg.LANG_TOKENS},{Cfg.LANG_DIM}), got {tuple(tok.shape)}")
        am = torch.as_tensor(np.asarray(attention_mask) if not torch.is_tensor(attention_mask) else attention_mask)
        am = am.to(dev).to(torch.int32).contiguous()
        cls = torch.as_tensor(np.ascontiguousarray(init_cls, dtype=np.float32) if not torch.is_tensor(init_cls) else init_cls)
        cls = cls.to(dev, torch.float32).contiguous()
        if tuple(am.shape) != (T, Cfg.LANG_TOKENS) or tuple(cls.shape) != (T, Cfg.DINO_DIM):
            raise ValueError("attention_mask must be (T,32) and the initial-image CLS embedding (T,768)")
        pad_ptr = None
        if lang_pad is not None:
            pad = torch.as_tensor(np.asarray(lang_pad)).to(dev).to(torch.uint8).contiguous()
            pad_ptr = pad.data_ptr()
        f16 = self.hn_blob_f16.data_ptr() if self.hn_blob_f16 is not None else None
        if rows is None:
            out = torch.empty((T, self.ngp), dtype=self.tdtype, device=dev)
            ctx = torch.empty((T, Cfg.CTX_DIM), dtype=torch.float32, device=dev)
            with self._on_device():
                ws, ws_bytes = self.workspace(0, T)
                if self.spec == M.MIX:
                    st = self.lib.hvla_generate(self.stream(), self.hn_blob.data_ptr(), f16, self.heads_w.data_ptr(), self.heads_b.data_ptr(),
                                                tok.data_ptr(), am.data_ptr(), pad_ptr, cls.data_ptr(), T, out.data_ptr(), ctx.data_ptr(),
                                                ws, ws_bytes, self.dtype)
                else:       # another generated-row layout (discrete head): same kernels, explicit row stride
                    st = self.lib.hvla_generate_n(self.stream(), self.hn_blob.data_ptr(), f16, self.heads_w.data_ptr(), self.heads_b.data_ptr(),
                                                  tok.data_ptr(), am.data_ptr(), pad_ptr, cls.data_ptr(), T, self.ngp, out.data_ptr(),
                                                  ctx.data_ptr(), ws, ws_bytes, self.dtype)
            N.check(st, "hvla_generate")
            return out, ctx
        if self.spec != M.MIX:
            raise ValueError("in-place row regeneration is implemented for the mix-head layout")
        # task-switch scheduler: regenerate rows `rows` of the persistent (weights, ctx) buffers in place
        out, ctx = into
        T_max = int(out.shape[0])
        ridx = np.asarray(rows.cpu() if torch.is_tensor(rows) else rows).astype(np.int64).ravel()
        if ridx.shape != (T,) or (T and (ridx.min() < 0 or ridx.max() >= T_max)) or len(set(ridx.tolist())) != T:
            raise ValueError(f"task_ids must be {T} distinct slots in [0, {T_max})")
        if out.dtype != self.tdtype or tuple(out.shape[1:]) != (M.N_GENERATED_PADDED,) or not out.is_contiguous():
            raise ValueError("persistent weight buffer does not match this runtime")
        rdev = torch.from_numpy(ridx.astype(np.int32)).to(dev)
        with self._on_device():
            ws, ws_bytes = self.workspace(0, T)
            st = self.lib.hvla_generate_rows(self.stream(), self.hn_blob.data_ptr(), f16, self.heads_w.data_ptr(), self.heads_b.data_ptr(),
                                             tok.data_ptr(), am.data_ptr(), pad_ptr, cls.data_ptr(), T, rdev.data_ptr(), T_max,
                                             out.data_ptr(), ctx.data_ptr() if ctx is not None else None, ws, ws_bytes, self.dtype)
        N.check(st, "hvla_generate_rows")
        return out, ctx

    # ---- act: one CUDA graph (~110 kernels) per (batch, weights, task map), replayed every control step ------------
    def _tidx(self, task_index, B, T):
        torch = _torch()
        if task_index is None:
            if T not in (1, B):
                raise ValueError(f"task_index is required when T ({T}) is neither 1 nor B ({B})")
            return None, None
        ti = torch.as_tensor(np.asarray(task_index) if not torch.is_tensor(task_index) else task_index)
        if ti.is_cuda and (ti.dtype != torch.int32 or not ti.is_contiguous()):
            raise ValueError("a CUDA task_index must be a contiguous int32 tensor (it is read in place by the captured graph)")
        ti = ti.to(self.device).to(torch.int32).contiguous()
        if tuple(ti.shape) != (B,):
            raise ValueError("task_index must have shape (B,)")
        if int(ti.min()) < 0 or int(ti.max()) >= T:
            raise ValueError("task_index out of range")
        return ti, ti.data_ptr()

    def _act_eager(self, img_dev, weights, tptr, B, T, act_dev, logit_dev):
        with self._on_device():
            ws, ws_bytes = self.workspace(B, 0)
            N.check(self.lib.hvla_act(self.stream(), self.dino_vec.data_ptr(), self.dino_mat.data_ptr(), img_dev.data_ptr(),
                                      weights.data_ptr(), tptr, B, T, act_dev.data_ptr(), logit_dev.data_ptr(), ws, ws_bytes,
                                      self.dtype), "hvla_act")

    def _graph(self, B, weights, task_index):
        """Static device buffers + captured graph of hvla_act for this (B, weights, task map)."""
        torch = _torch()
        T = int(weights.shape[0])
        # key without a device sync: a CUDA task map is identified by its address (the graph reads its CONTENTS at replay, so
        # in-place edits of the map are picked up; its range is validated when the graph is built), a host map by its bytes
        if task_index is None:
            tkey = None
        elif torch.is_tensor(task_index) and task_index.is_cuda:
            tkey = ("dev", int(task_index.data_ptr()), int(task_index.numel()), str(task_index.dtype))
        else:
            tkey = np.asarray(task_index).tobytes()
        key = (B, int(weights.data_ptr()), tkey)
        st = self._graphs.get(key)
        if st is not None and st["ws_key"] == self._ws_key:
            return st
        if len(self._graphs) >= 8:
            self._graphs.pop(next(iter(self._graphs)))
        tidx_t, tptr = self._tidx(task_index, B, T)
        st = {
            "img_dev": torch.empty((B, Cfg.IMAGE_SIZE, Cfg.IMAGE_SIZE, 3), dtype=torch.uint8, device=self.device),
            "act_dev": torch.empty((B, Cfg.ACTION_HORIZON, Cfg.ACTION_DIM), dtype=torch.float32, device=self.device),
            "logit_dev": torch.empty((B, Cfg.ACTION_HORIZON), dtype=torch.float32, device=self.device),
            "weights": weights, "tidx": tidx_t, "tptr": tptr, "T": T, "graph": None,
        }
        self.workspace(B, 0)
        if self.use_graphs and not self._profiling:
            st["img_dev"].zero_()
            n0 = int(self.lib.hvla_launch_count())
            self._act_eager(st["img_dev"], weights, tptr, B, T, st["act_dev"], st["logit_dev"])   # one-time kernel setup outside capture
            st["n_launches"] = int(self.lib.hvla_launch_count()) - n0
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with self._on_device(), torch.cuda.graph(graph):
                self._act_eager(st["img_dev"], weights, tptr, B, T, st["act_dev"], st["logit_dev"])
            st["graph"] = graph
            self.graph_captures += 1
        st["ws_key"] = self._ws_key
        self._graphs[key] = st
        return st

    def _run(self, st, B):
        if st["graph"] is not None and not self._profiling:
            st["graph"].replay()
            self.replayed_launches += st["n_launches"]
        else:
            self._act_eager(st["img_dev"], st["weights"], st["tptr"], B, st["T"], st["act_dev"], st["logit_dev"])

    def act_device(self, images, weights, task_index=None):
        """images: uint8 CUDA tensor (B,224,224,3); returns (action, logit) CUDA tensors (asynchronous)."""
        torch = _torch()
        B = int(images.shape[0])
        if tuple(images.shape[1:]) != (Cfg.IMAGE_SIZE, Cfg.IMAGE_SIZE, 3) or images.dtype != torch.uint8:
            raise ValueError("Input image size must be 224x224 (uint8, NHWC)")   # base_vit.py:87-89
        if B == 0:
            return (torch.empty((0, Cfg.ACTION_HORIZON, Cfg.ACTION_DIM), dtype=torch.float32, device=self.device),
                    torch.empty((0, Cfg.ACTION_HORIZON), dtype=torch.float32, device=self.device))
        st = self._graph(B, weights, task_index)
        st["img_dev"].copy_(images, non_blocking=True)
        self._run(st, B)
        return st["act_dev"].clone(), st["logit_dev"].clone()

    def act_host(self, images, weights, task_index=None):
        """images: host uint8 (numpy or pinned torch CPU tensor) (B,224,224,3); returns numpy
        (action (B,4,7), logit (B,4)).  H2D copy + graph replay + D2H copy + one stream sync."""
        torch = _torch()
        if torch.is_tensor(images):
            if images.dtype != torch.uint8 or tuple(images.shape[1:]) != (Cfg.IMAGE_SIZE, Cfg.IMAGE_SIZE, 3):
                raise ValueError("Input image size must be 224x224 (uint8, NHWC)")
            src = images.contiguous()
        else:
            arr = np.ascontiguousarray(images)
            if arr.dtype != np.uint8 or arr.shape[1:] != (Cfg.IMAGE_SIZE, Cfg.IMAGE_SIZE, 3):
                raise ValueError("Input image size must be 224x224 (uint8, NHWC)")
            src = self.pinned("img", arr.shape, torch.uint8)
            if arr.size:
                src.numpy()[...] = arr
        B = int(src.shape[0])
        if B == 0:
            return (np.zeros((0, Cfg.ACTION_HORIZON, Cfg.ACTION_DIM), np.float32), np.zeros((0, Cfg.ACTION_HORIZON), np.float32))
        st = self._graph(B, weights, task_index)
        act = self.pinned("act", (B, Cfg.ACTION_HORIZON, Cfg.ACTION_DIM), torch.float32)
        logit = self.pinned("logit", (B, Cfg.ACTION_HORIZON), torch.float32)
        st["img_dev"].copy_(src, non_blocking=True)
        self._run(st, B)
        act.copy_(st["act_dev"], non_blocking=True)
        logit.copy_(st["logit_dev"], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return act.numpy().copy(), logit.numpy().copy()

    def act_discrete(self, images, weights, task_index=None, want_top2: bool = False):
        """DiscreteActionHead variant of the act step (hvla_act_discrete): DINOv2 -> base ViT with 4 / 28 readout tokens ->
        vocab_proj -> argmax -> BinTokenizer.decode.  -> (action (B,4,7) f32 bin centres, tokens (B,4,7) i32[, top2 (B,4,7,2)]), CUDA tensors."""
        torch = _torch()
        img = images if torch.is_tensor(images) else torch.from_numpy(np.ascontiguousarray(images))
        img = img.to(self.device).contiguous()
        B, T = int(img.shape[0]), int(weights.shape[0])
        if tuple(img.shape[1:]) != (Cfg.IMAGE_SIZE, Cfg.IMAGE_SIZE, 3) or img.dtype != torch.uint8:
            raise ValueError("Input image size must be 224x224 (uint8, NHWC)")
        if self.spec.kind != "discrete" or int(weights.shape[1]) != self.ngp:
            raise ValueError("act_discrete needs a model configured with action_head_type='discrete' and its generated weights")
        keep, tptr = self._tidx(task_index, B, T)
        act = torch.empty((B, Cfg.ACTION_HORIZON, Cfg.ACTION_DIM), dtype=torch.float32, device=self.device)
        tok = torch.empty((B, Cfg.ACTION_HORIZON, Cfg.ACTION_DIM), dtype=torch.int32, device=self.device)
        top2 = torch.empty((B, Cfg.ACTION_HORIZON, Cfg.ACTION_DIM, 2), dtype=torch.float32, device=self.device) if want_top2 else None
        if B:
            with self._on_device():
                ws, ws_bytes = self.workspace(B, 0)
                N.check(self.lib.hvla_act_discrete(self.stream(), self.dino_vec.data_ptr(), self.dino_mat.data_ptr(), img.data_ptr(),
                                                   weights.data_ptr(), tptr, B, T, self.spec.n_action_tokens, act.data_ptr(), tok.data_ptr(),
                                                   top2.data_ptr() if top2 is not None else None, ws, ws_bytes, self.dtype), "hvla_act_discrete")
        return (act, tok, top2) if want_top2 else (act, tok)

    def act_debug(self, images, weights, task_index=None):
        """One eager act step that also returns the attention weights the reference sows as ``intermediates``
        (hvla_act_debug): -> (action (B,4,7), logit (B,4), dino_maps (12,B,12,257,257), base_maps (4,B,4,257,257)), CUDA
        tensors.  Debugging call (38 MB of maps per image, base net on the generic kernels)."""
        torch = _torch()
        img = images if torch.is_tensor(images) else torch.from_numpy(np.ascontiguousarray(images))
        img = img.to(self.device).contiguous()
        B, T = int(img.shape[0]), int(weights.shape[0])
        if tuple(img.shape[1:]) != (Cfg.IMAGE_SIZE, Cfg.IMAGE_SIZE, 3) or img.dtype != torch.uint8:
            raise ValueError("Input image size must be 224x224 (uint8, NHWC)")
        keep, tptr = self._tidx(task_index, B, T)
        S = Cfg.DINO_TOKENS
        act = torch.empty((B, Cfg.ACTION_HORIZON, Cfg.ACTION_DIM), dtype=torch.float32, device=self.device)
        logit = torch.empty((B, Cfg.ACTION_HORIZON), dtype=torch.float32, device=self.device)
        dmaps = torch.empty((Cfg.DINO_LAYERS, B, Cfg.DINO_HEADS, S, S), dtype=torch.float32, device=self.device)
        bmaps = torch.empty((Cfg.BASE_LAYERS, B, Cfg.BASE_HEADS, S, S), dtype=torch.float32, device=self.device)
        if B:
            with self._on_device():
                ws, ws_bytes = self.workspace(B, 0)
                N.check(self.lib.hvla_act_debug(self.stream(), self.dino_vec.data_ptr(), self.dino_mat.data_ptr(), img.data_ptr(),
                                                weights.data_ptr(), tptr, B, T, act.data_ptr(), logit.data_ptr(), dmaps.data_ptr(),
                                                bmaps.data_ptr(), ws, ws_bytes, self.dtype), "hvla_act_debug")
        return act, logit, dmaps, bmaps

    def profile(self, fn, repeats: int = 1) -> dict:
        """Run ``fn`` with per-kernel-class event timing on; -> {class: (launches, total_ms)} per repeat."""
        self.lib.hvla_profile_enable(1)
        self._profiling = True
        try:
            for _ in range(repeats):
                fn()
            buf = C.create_string_buffer(4096)
            N.check(self.lib.hvla_profile_report(buf, 4096), "hvla_profile_report")
        finally:
            self._profiling = False
            self.lib.hvla_profile_enable(0)
        out = {}
        for line in buf.value.decode().splitlines():
            name, n, ms = line.split()
            out[name] = (int(n) / repeats, float(ms) / repeats)
        return out

    def pack_dino_tree(self, tree: dict):
        """Upload another DINOv2-base param tree (e.g. the frozen pretrained encoder of the initial image,
        data/simpler/evaluate.py:146-163) -> (vec, mat) device blobs for ``dino_forward(..., blobs=...)``."""
        torch = _torch()
        return self._dino_blobs(*P.pack_dino_tree(tree, transposed=(self.precision == "bf16")))

    def dino_forward(self, images, blobs=None):
        """images uint8 CUDA (B,224,224,3) -> last_hidden_state (B,257,768) in the runtime dtype."""
        torch = _torch()
        B = int(images.shape[0])
        if tuple(images.shape[1:]) != (Cfg.IMAGE_SIZE, Cfg.IMAGE_SIZE, 3) or images.dtype != torch.uint8:
            raise ValueError("Input image size must be 224x224 (uint8, NHWC)")
        vec, mat = blobs if blobs is not None else (self.dino_vec, self.dino_mat)
        out = torch.empty((B, Cfg.DINO_TOKENS, Cfg.DINO_DIM), dtype=self.tdtype, device=self.device)
        images = images.contiguous()
        with self._on_device():
            ws, ws_bytes = self.workspace(B, 0)
            st = self.lib.hvla_dino_forward(self.stream(), vec.data_ptr(), mat.data_ptr(),
                                            images.data_ptr(), B, out.data_ptr(), ws, ws_bytes, self.dtype)
        N.check(st, "hvla_dino_forward")
        return out

    def base_act(self, emb, weights, task_index=None):
        torch = _torch()
        B, T = int(emb.shape[0]), int(weights.shape[0])
        keep, tptr = self._tidx(task_index, B, T)
        act = torch.empty((B, Cfg.ACTION_HORIZON, Cfg.ACTION_DIM), dtype=torch.float32, device=self.device)
        logit = torch.empty((B, Cfg.ACTION_HORIZON), dtype=torch.float32, device=self.device)
        emb = emb.contiguous()
        with self._on_device():
            ws, ws_bytes = self.workspace(B, 0)
            st = self.lib.hvla_base_act(self.stream(), emb.data_ptr(), weights.data_ptr(), tptr, B, T,
                                        act.data_ptr(), logit.data_ptr(), ws, ws_bytes, self.dtype)
        N.check(st, "hvla_base_act")
        return act, logit
