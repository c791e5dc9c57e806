"""Python owner of the flat device buffers + thin wrapper over the C ABI (include/hippie_b200.h).

PyTorch is plumbing here: it allocates device memory, provides the CUDA stream and (for data
parallel training) the NCCL all-reduce of the flat gradient buffer.  All arithmetic of the
cVAE step happens inside libhippie_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch

from . import _lib

LAYOUT_NATIVE = 0
LAYOUT_CONV_OKI = 1
# hippie_cfg.multimodal (include/hippie_b200.h: HIPPIE_KIND_*)
KIND_UNIMODAL, KIND_MULTIMODAL, KIND_ENCODER, KIND_DECODER = 0, 1, 2, 3
FLAG_SOURCE_LABEL, FLAG_CLASS_LABEL, FLAG_PAIR_SATURATED, FLAG_WEIGHT_SATURATED = 1, 2, 4, 8


@dataclass(frozen=True)
class ParamInfo:
    name: str
    offset: int
    numel: int
    shape: Tuple[int, ...]
    layout: int


@dataclass(frozen=True)
class BnInfo:
    name: str
    offset: int
    channels: int
    index: int


@dataclass(frozen=True)
class TensorInfo:
    name: str
    offset: int
    L: int
    C: int
    pad: int


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class Engine:
    """One handle per (process, device).  `Engine(...)` only builds the layout (no CUDA needed);
    `allocate(device)` creates the flat buffers on the GPU and binds them."""

    def __init__(self, z_dim: int, len_wave: int = 50, len_isi: int = 100, class_hidden_dim: int = 5,
                 num_sources: int = 5, num_classes: int = 5, multimodal: bool = True, max_batch: int = 512,
                 inference_only: bool = False, conv_path: int = 0):
        self._L = _lib.lib()
        self.kind = int(multimodal)  # bool (uni / multimodal cVAE) or KIND_ENCODER / KIND_DECODER (stand-alone backbone)
        self.cfg = _lib.HippieCfg(z_dim, class_hidden_dim, num_sources, num_classes, len_wave, len_isi,
                                  self.kind, max_batch, 1 if inference_only else 0, conv_path)
        self._h = C.c_void_p()
        rc = self._L.hippie_create(C.byref(self.cfg), C.byref(self._h))
        if rc != 0:
            raise ValueError(f"hippie_create failed ({rc}): invalid configuration")
        self.inference_only = bool(inference_only) or self.kind in (KIND_ENCODER, KIND_DECODER)
        self.z_dim, self.multimodal, self.max_batch = z_dim, self.kind == KIND_MULTIMODAL, max_batch
        self.class_hidden_dim, self.num_sources, self.num_classes = class_hidden_dim, num_sources, num_classes
        self.len_wave, self.len_isi = len_wave, len_isi
        self.param_floats = int(self._L.hippie_param_floats(self._h))
        self.bn_floats = int(self._L.hippie_bn_floats(self._h))
        self.workspace_bytes = int(self._L.hippie_workspace_bytes(self._h))
        self.params: List[ParamInfo] = []
        name = C.create_string_buffer(256)
        off, numel, ndim, layout = C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32()
        shape = (C.c_int64 * 3)()
        for i in range(self._L.hippie_num_params(self._h)):
            self._L.hippie_param_info(self._h, i, name, 256, C.byref(off), C.byref(numel), C.byref(ndim), shape,
                                      C.byref(layout))
            self.params.append(ParamInfo(name.value.decode(), off.value, numel.value,
                                         tuple(int(shape[k]) for k in range(ndim.value)), layout.value))
        self.bns: List[BnInfo] = []
        ch = C.c_int64()
        for i in range(self._L.hippie_num_bn(self._h)):
            self._L.hippie_bn_info(self._h, i, name, 256, C.byref(off), C.byref(ch))
            self.bns.append(BnInfo(name.value.decode(), off.value, ch.value, i))
        self.tensors: List[TensorInfo] = []
        tl, tc, tp = C.c_int32(), C.c_int32(), C.c_int32()
        for i in range(self._L.hippie_num_tensors(self._h)):
            self._L.hippie_tensor_info(self._h, i, name, 256, C.byref(off), C.byref(tl), C.byref(tc), C.byref(tp))
            if off.value >= 0:  # tensors that only exist as fp16 pair planes (GEMM operands) have no fp32 view
                self.tensors.append(TensorInfo(name.value.decode(), off.value, tl.value, tc.value, tp.value))
        self.device: Optional[torch.device] = None
        self.flat_params = self.flat_grads = self.exp_avg = self.exp_avg_sq = None
        self._pver = -1  # version counter of flat_params the engine's weight planes correspond to
        self.bn_mean = self.bn_var = self.bn_count = self.workspace = None

    # ------------------------------------------------------------------------------------------
    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                self._L.hippie_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            msg = self._L.hippie_last_error(self._h)
            raise RuntimeError(f"libhippie_b200 error {rc}: {msg.decode() if msg else ''}")

    @staticmethod
    def _stream() -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    # ------------------------------------------------------------------------------------------
    def allocate(self, device="cuda:0", guard: int = 0):
        """Creates the flat fp32 buffers (params | grads | exp_avg | exp_avg_sq), BatchNorm buffers and the
        workspace on `device` and hands them to the engine (hippie_bind).

        `guard` > 0 (tests): every buffer is carved out of a larger allocation with `guard` canary elements on either
        side; `check_guards()` verifies afterwards that no kernel wrote outside the buffers it was given (the pool this
        repository is developed on does not allow compute-sanitizer, tests/test_gpu_parity2.py)."""
        device = torch.device(device)
        if device.type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("hippie_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        self.device = device
        self._guards = []

        def buf(n, dtype, fill):
            if not guard:
                return torch.full((n,), fill, dtype=dtype, device=device)
            canary = {torch.float32: float("nan"), torch.int64: -0x5A5A5A5A5A5A5A5B, torch.uint8: 0xA5}[dtype]
            whole = torch.full((n + 2 * guard,), canary, dtype=dtype, device=device)
            whole[guard:guard + n] = fill
            self._guards.append((whole, guard, n))
            return whole[guard:guard + n]

        with torch.cuda.device(device):
            n = self.param_floats
            self.flat_params = buf(n, torch.float32, 0)
            if not self.inference_only:
                self.flat_grads = buf(n, torch.float32, 0)
                self.exp_avg = buf(n, torch.float32, 0)
                self.exp_avg_sq = buf(n, torch.float32, 0)
            self.bn_mean = buf(self.bn_floats, torch.float32, 0)
            self.bn_var = buf(self.bn_floats, torch.float32, 1)
            self.bn_count = buf(len(self.bns), torch.int64, 0)
            self.workspace = buf(self.workspace_bytes, torch.uint8, 0) if guard else \
                torch.empty(self.workspace_bytes, dtype=torch.uint8, device=device)
            self._check(self._L.hippie_bind(self._h, _ptr(self.flat_params), _ptr(self.flat_grads), _ptr(self.exp_avg),
                                            _ptr(self.exp_avg_sq), _ptr(self.bn_mean), _ptr(self.bn_var),
                                            _ptr(self.bn_count), _ptr(self.workspace), self.workspace_bytes,
                                            self._stream()))
        return self

    def check_guards(self):
        """Raises if any canary element around the bound buffers changed (see `allocate(guard=...)`)."""
        torch.cuda.synchronize(self.device)
        for whole, g, n in self._guards:
            for side, part in (("before", whole[:g]), ("after", whole[g + n:])):
                ref = torch.full_like(part, float("nan") if whole.dtype == torch.float32 else
                                      (-0x5A5A5A5A5A5A5A5B if whole.dtype == torch.int64 else 0xA5))
                a = part.view(torch.int32) if whole.dtype == torch.float32 else part
                b = ref.view(torch.int32) if whole.dtype == torch.float32 else ref
                if not torch.equal(a, b):
                    bad = int((a != b).sum())
                    raise AssertionError(f"{bad} canary elements {side} a {whole.dtype} buffer of {n} elements were overwritten")

    # ---- views ---------------------------------------------------------------------------------
    @staticmethod
    def view_of(flat: torch.Tensor, p: ParamInfo) -> torch.Tensor:
        """The torch-shaped view of one parameter inside a flat buffer.  Conv1d weights are stored
        [Cout][k][Cin] (channels-last GEMM operand); the view is permuted back to torch's [Cout][Cin][k]."""
        t = flat[p.offset:p.offset + p.numel]
        if p.layout == LAYOUT_CONV_OKI:
            co, ci, k = p.shape
            return t.view(co, k, ci).permute(0, 2, 1)
        return t.view(p.shape)

    def tensor_view(self, name: str, B: int) -> torch.Tensor:
        """[B, C, L] view (torch layout) of a named activation / gradient tensor in the workspace."""
        ti = next(t for t in self.tensors if t.name == name)
        ws = self.workspace.view(torch.float32)
        rows = ti.L + 2 * ti.pad
        t = ws[ti.offset:ti.offset + B * rows * ti.C].view(B, rows, ti.C)
        return t[:, ti.pad:ti.pad + ti.L, :].permute(0, 2, 1)

    # ---- calls ---------------------------------------------------------------------------------
    def params_changed(self):
        """Tells the engine that `flat_params` was written by something other than `clip_adamw` (hippie_params_changed):
        the tensor-core path converts the whole buffer into its fp16 pair planes again at the next call.  In-place torch
        operations on `flat_params` or on any view of it (nn.Parameter.copy_, load_state_dict, init functions, a
        broadcast) are detected through the tensor's version counter; writes that bypass it (`param.data.<op>_()`, raw
        pointers) need this call."""
        self._check(self._L.hippie_params_changed(self._h))
        self._pver = self.flat_params._version

    def _sync_params(self):
        if self.flat_params._version != self._pver:
            self.params_changed()

    def _io(self, x1, x2, src, cls, eps, B):
        self._sync_params()
        assert x1.dtype == torch.float32 and x1.is_contiguous() and x1.is_cuda
        assert src.dtype == torch.int64 and src.is_contiguous()
        if self.multimodal:
            assert x2 is not None and x2.dtype == torch.float32 and x2.is_contiguous()
            assert x2.numel() == B * self.len_isi
        assert x1.numel() == B * self.len_wave
        if cls is not None:
            assert cls.dtype == torch.int64 and cls.is_contiguous()
        if eps is not None:
            assert eps.dtype == torch.float32 and eps.is_contiguous() and eps.numel() == B * self.z_dim

    def train_fwd_bwd(self, x1, x2, src, cls, eps, beta: float, w1: float = 1.0, w2: float = 1.0, scalars=None,
                      outputs: bool = False):
        """training_step + zero_grad + backward.  Returns (scalars[8] device tensor, outputs dict or None)."""
        B = x1.shape[0]
        self._io(x1, x2, src, cls, eps, B)
        if scalars is None:
            scalars = torch.zeros(8, dtype=torch.float32, device=self.device)
        outs = self._out_buffers(B) if outputs else {}
        self._check(self._L.hippie_train_fwd_bwd(
            self._h, _ptr(x1), _ptr(x2), _ptr(src), _ptr(cls), _ptr(eps), B, beta, w1, w2, _ptr(scalars),
            _ptr(outs.get("enc")), _ptr(outs.get("mu")), _ptr(outs.get("logvar")), _ptr(outs.get("dec1")),
            _ptr(outs.get("dec2")), self._stream()))
        return scalars, (outs if outputs else None)

    def train_fwd_bwd_part(self, part: int, x1, x2, src, cls, eps, beta: float, w1: float = 1.0, w2: float = 1.0,
                           scalars=None):
        """Part 0 (forward, loss, decoder + head backward) or part 1 (encoder backward) of the train step; after part 0
        `flat_grads[grad_split:]` is final, after part 1 `flat_grads[:grad_split]`; parts 2 + 3 replace part 1 with the
        deep / shallow halves of the encoders (`grad_bounds`) (hippie_train_fwd_bwd_part)."""
        B = x1.shape[0]
        self._io(x1, x2, src, cls, eps, B)
        if scalars is None and part in (0, 4):
            scalars = torch.zeros(8, dtype=torch.float32, device=self.device)
        self._check(self._L.hippie_train_fwd_bwd_part(self._h, _ptr(x1), _ptr(x2), _ptr(src), _ptr(cls), _ptr(eps), B, beta,
                                                      w1, w2, _ptr(scalars), part, self._stream()))
        return scalars

    def slice_wait(self, k: int, stream: "torch.cuda.Stream"):
        """Makes `stream` wait until slice k of the gradient buffer of the most recent part-4 step is final
        (0: flat_grads[grad_split:], 1: the deep halves of `grad_bounds`); hippie_slice_wait."""
        self._check(self._L.hippie_slice_wait(self._h, k, C.c_void_p(stream.cuda_stream)))

    @property
    def grad_split(self) -> int:
        return int(self._L.hippie_grad_split(self._h))

    @property
    def grad_bounds(self):
        """[(begin, deep, end)] per encoder: flat_grads[deep:end] is final after part 2, flat_grads[begin:deep] after
        part 3 (hippie_grad_bounds); empty ranges are dropped."""
        import ctypes as C
        b = (C.c_int64 * 6)()
        self._check(self._L.hippie_grad_bounds(self._h, b))
        return [tuple(int(v) for v in b[3 * e:3 * e + 3]) for e in range(2) if b[3 * e + 2] > b[3 * e]]

    def train_forward(self, x1, x2, src, cls, eps, beta: float = 1.0, w1: float = 1.0, w2: float = 1.0, scalars=None):
        return self.eval_forward(x1, x2, src, cls, eps, beta, w1, w2, scalars, _fn="hippie_train_forward")

    def eval_forward(self, x1, x2, src, cls, eps, beta: float = 1.0, w1: float = 1.0, w2: float = 1.0, scalars=None,
                     _fn="hippie_eval_forward"):
        B = x1.shape[0]
        self._io(x1, x2, src, cls, eps, B)
        outs = self._out_buffers(B)
        self._check(getattr(self._L, _fn)(
            self._h, _ptr(x1), _ptr(x2), _ptr(src), _ptr(cls), _ptr(eps), B, beta, w1, w2, _ptr(scalars),
            _ptr(outs.get("enc")), _ptr(outs.get("mu")), _ptr(outs.get("logvar")), _ptr(outs.get("dec1")),
            _ptr(outs.get("dec2")), self._stream()))
        return outs

    def embed(self, x1, x2, src, cls=None, zscore_ddof: int = -1, out=None):
        B = x1.shape[0]
        self._io(x1, x2, src, cls, None, B)
        z = self.z_dim
        if out is None:
            out = {k: torch.empty(B, z, dtype=torch.float32, device=self.device) for k in ("enc", "mu", "logvar")}
        self._check(self._L.hippie_embed(self._h, _ptr(x1), _ptr(x2), _ptr(src), _ptr(cls), B, zscore_ddof,
                                         _ptr(out["enc"]), _ptr(out["mu"]), _ptr(out["logvar"]), self._stream()))
        return out

    def clip_adamw(self, lr: float, weight_decay: float, step: int, max_norm: Optional[float] = None,
                   grad_scale: float = 1.0, betas=(0.9, 0.999), eps: float = 1e-8, step_cls: int = 0,
                   has_cls_grad: bool = False, scalars=None):
        if scalars is None:
            scalars = torch.zeros(8, dtype=torch.float32, device=self.device)
        self._check(self._L.hippie_clip_adamw(self._h, lr, betas[0], betas[1], eps, weight_decay,
                                              -1.0 if max_norm is None else float(max_norm), grad_scale, step, step_cls,
                                              1 if has_cls_grad else 0, _ptr(scalars), self._stream()))
        return scalars

    # ---- module-level forward calls (hippie/backbones.py forward(), MultiModalCVAE.encode / .decode) -------------
    def encoder_forward(self, x, which: int = 0, train: bool = False):
        B = x.shape[0]
        self._sync_params()
        assert x.dtype == torch.float32 and x.is_contiguous() and x.is_cuda
        out = torch.empty(B, 2 * self.z_dim, dtype=torch.float32, device=self.device)
        self._check(self._L.hippie_encoder_forward(self._h, which, _ptr(x), B, 1 if train else 0, _ptr(out), self._stream()))
        return out

    def decoder_forward(self, d, which: int = 0, train: bool = False):
        B = d.shape[0]
        self._sync_params()
        assert d.dtype == torch.float32 and d.is_contiguous() and d.is_cuda and d.numel() == B * 2 * self.z_dim
        L = self.len_isi if (self.multimodal and which == 1) else self.len_wave
        out = torch.empty(B, 1, L, dtype=torch.float32, device=self.device)
        self._check(self._L.hippie_decoder_forward(self._h, which, _ptr(d), B, 1 if train else 0, _ptr(out), self._stream()))
        return out

    def encode(self, x1, x2, source_emb, class_emb, train: bool = False):
        B = x1.shape[0]
        self._sync_params()
        for t in (x1, x2, source_emb, class_emb):
            assert t is None or (t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda)
        assert source_emb.shape == class_emb.shape == (B, self.class_hidden_dim)
        o = {k: torch.empty(B, self.z_dim, dtype=torch.float32, device=self.device) for k in ("enc", "mu", "logvar")}
        self._check(self._L.hippie_encode(self._h, _ptr(x1), _ptr(x2), _ptr(source_emb), _ptr(class_emb), B, 1 if train else 0,
                                          _ptr(o["enc"]), _ptr(o["mu"]), _ptr(o["logvar"]), self._stream()))
        return o["enc"], o["mu"], o["logvar"]

    def decode(self, z, source_emb, class_emb, train: bool = False):
        B = z.shape[0]
        self._sync_params()
        for t in (z, source_emb, class_emb):
            assert t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda
        assert z.shape == (B, self.z_dim) and source_emb.shape == class_emb.shape == (B, self.class_hidden_dim)
        d1 = torch.empty(B, 1, self.len_wave, dtype=torch.float32, device=self.device)
        d2 = torch.empty(B, 1, self.len_isi, dtype=torch.float32, device=self.device) if self.multimodal else None
        self._check(self._L.hippie_decode(self._h, _ptr(z), _ptr(source_emb), _ptr(class_emb), B, 1 if train else 0, _ptr(d1),
                                          _ptr(d2), self._stream()))
        return d1, d2

    def device_flags(self, clear: bool = True) -> int:
        """HIPPIE_FLAG_* word set by the kernels (bad label indices, fp16 pair-plane saturation).  Synchronises."""
        f = C.c_uint32(0)
        self._check(self._L.hippie_device_flags(self._h, C.byref(f), 1 if clear else 0, self._stream()))
        return int(f.value)

    def raise_on_flags(self):
        """What the reference turns into exceptions: nn.Embedding's IndexError for a label outside the table; fp16
        pair-plane saturation has no counterpart in the reference (fp32 operands) and raises OverflowError."""
        f = self.device_flags()
        if f & FLAG_SOURCE_LABEL:
            raise IndexError(f"index out of range in self: a source label lies outside [0, {self.num_sources})")
        if f & FLAG_CLASS_LABEL:
            raise IndexError(f"index out of range in self: a class label lies outside [0, {self.num_classes})")
        if f & (FLAG_PAIR_SATURATED | FLAG_WEIGHT_SATURATED):
            what = "a parameter exceeded |w| < 255.9" if f & FLAG_WEIGHT_SATURATED else "an activation exceeded 65504"
            raise OverflowError("hippie_b200: " + what + ": outside the range of the fp16 pair planes of the tensor-core "
                                "GEMMs (use conv_path=1, the FP32 CUDA-core path, for such models)")

    def conv_path_in_use(self) -> int:
        """2 = tcgen05 implicit GEMMs over fp16 pair planes, 1 = FP32 CUDA-core GEMMs (valid after allocate())."""
        return int(self._L.hippie_conv_path_in_use(self._h))

    def last_launch_count(self) -> int:
        return int(self._L.hippie_last_launch_count(self._h))

    def _out_buffers(self, B):
        z, dev = self.z_dim, self.device
        o = {k: torch.empty(B, z, dtype=torch.float32, device=dev) for k in ("enc", "mu", "logvar")}
        o["dec1"] = torch.empty(B, 1, self.len_wave, dtype=torch.float32, device=dev)
        if self.multimodal:
            o["dec2"] = torch.empty(B, 1, self.len_isi, dtype=torch.float32, device=dev)
        return o

    # ---- state exchange (used by the nn.Module mirror and the tests) ---------------------------
    def load_named(self, state: dict, strict: bool = True):
        """Copies a {reference state_dict key: tensor} mapping into the flat buffers."""
        missing = []
        with torch.no_grad():
            for p in self.params:
                if p.name in state:
                    self.view_of(self.flat_params, p).copy_(state[p.name].to(self.device, torch.float32))
                else:
                    missing.append(p.name)
            for b in self.bns:
                for key, buf in ((".running_mean", self.bn_mean), (".running_var", self.bn_var)):
                    if b.name + key in state:
                        buf[b.offset:b.offset + b.channels].copy_(state[b.name + key].to(self.device, torch.float32))
                    else:
                        missing.append(b.name + key)
                if b.name + ".num_batches_tracked" in state:
                    self.bn_count[b.index] = int(state[b.name + ".num_batches_tracked"])
        if strict and missing:
            raise KeyError(f"missing keys: {missing[:5]}{'...' if len(missing) > 5 else ''}")
        return missing

    def named_state(self) -> dict:
        out = {}
        for p in self.params:
            out[p.name] = self.view_of(self.flat_params, p)
        for b in self.bns:
            out[b.name + ".running_mean"] = self.bn_mean[b.offset:b.offset + b.channels]
            out[b.name + ".running_var"] = self.bn_var[b.offset:b.offset + b.channels]
            out[b.name + ".num_batches_tracked"] = self.bn_count[b.index]
        return out

    def named_grads(self) -> dict:
        return {p.name: self.view_of(self.flat_grads, p) for p in self.params}
