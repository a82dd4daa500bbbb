"""Host-side plumbing shared by the module facades: precision mode, device workspaces, weight cache.

PyTorch is used here only for device memory, streams and dtype bookkeeping; all arithmetic of the hot path happens
inside libhsenet_sm100a.so.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Dict, Iterable, Tuple

import torch

from . import _lib

_PRECISIONS = {"bf16": _lib.PREC_BF16, "fp32_verify": _lib.PREC_FP32_VERIFY}
_state = threading.local()
_default_precision = os.environ.get("HSENET_B200_PRECISION", "bf16")


def set_precision(name: str) -> None:
    """'bf16' (tcgen05 tensor-core kernels, the product path) or 'fp32_verify' (fp32 CUDA-core verification mode)."""
    if name not in _PRECISIONS:
        raise ValueError(f"unknown precision {name!r}; expected one of {sorted(_PRECISIONS)}")
    global _default_precision
    _default_precision = name


def get_precision() -> str:
    return getattr(_state, "override", None) or _default_precision


class precision:
    """Context manager: ``with hsenet_b200.precision('fp32_verify'): ...``"""

    def __init__(self, name: str):
        if name not in _PRECISIONS:
            raise ValueError(f"unknown precision {name!r}")
        self.name = name

    def __enter__(self):
        self.prev = getattr(_state, "override", None)
        _state.override = self.name
        return self

    def __exit__(self, *exc):
        _state.override = self.prev
        return False


def precision_code(name: str | None = None) -> int:
    return _PRECISIONS[name or get_precision()]


def act_dtype(name: str | None = None) -> torch.dtype:
    return torch.bfloat16 if (name or get_precision()) == "bf16" else torch.float32


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return _lib.DTYPE_F32
    if dt == torch.bfloat16:
        return _lib.DTYPE_BF16
    if dt == torch.float16:
        return _lib.DTYPE_F16
    raise ValueError(f"unsupported dtype {dt}")


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


_side_streams: dict = {}


def side_stream(device: torch.device) -> "torch.cuda.Stream":
    """One extra stream per device for work that runs beside the caller's stream (the second encoder of the tower)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    st = _side_streams.get(idx)
    if st is None:
        st = _side_streams[idx] = torch.cuda.Stream(device=device)
    return st


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"hsenet_b200: {what} must live on a CUDA device (got {t.device}); there is no CPU path. "
            "Move the module and its inputs to cuda.")


def forbid_autograd(params: Iterable[torch.Tensor], what: str) -> None:
    """Guard of the inference-only launch paths (raw-pointer writes, graph replays): they are invisible to autograd, so
    reaching one with trainable parameters under grad mode is a bug -- the modules route such calls to the training path
    (hsenet_b200/training.py) before getting here."""
    if torch.is_grad_enabled() and any(p.requires_grad for p in params):
        raise NotImplementedError(
            f"hsenet_b200.{what}: this launch path does not record an autograd graph; call the module's forward "
            "(training path) or run under torch.no_grad().")


# ---- workspaces: one growing buffer per (device, stream, tag) -----------------------------------------------------
# Keyed on the stream the call is enqueued on (re-entrant per (device, stream)): two streams that
# run the same stage concurrently never share scratch memory, and calls on one stream reuse theirs in stream order.
_workspaces: Dict[Tuple[int, int, str], torch.Tensor] = {}
_ws_lock = threading.Lock()


def workspace(device: torch.device, nbytes: int, tag: str) -> torch.Tensor:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, int(torch.cuda.current_stream(device).cuda_stream), tag)
    with _ws_lock:
        return _workspace_locked(key, device, nbytes)


def _workspace_locked(key, device, nbytes):
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _workspaces.pop(key, None)
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def release_workspaces() -> None:
    _workspaces.clear()


# ---- weight cache ----------------------------------------------------------------------------------------------------
def _sig(params) -> tuple:
    return tuple((p.data_ptr(), p._version, p.dtype, p.device) for p in params)


def fold_layernorm(norm, linear):
    """(w_folded bf16 [N,K], colsum f32 [N], bias_folded f32 [N]) of ``linear(norm(x))`` -- hsenet_fold_layernorm."""
    from . import _lib
    w = f32(linear.weight)
    n, k = w.shape
    dev = w.device
    wf = torch.empty(n, k, dtype=torch.bfloat16, device=dev)
    cs = torch.empty(n, dtype=torch.float32, device=dev)
    bf = torch.empty(n, dtype=torch.float32, device=dev)
    bias = None if linear.bias is None else f32(linear.bias)
    g, b = f32(norm.weight), f32(norm.bias)
    with torch.cuda.device(dev):
        rc = _lib.load().hsenet_fold_layernorm(w.data_ptr(), g.data_ptr(), b.data_ptr(), ptr(bias), n, k,
                                               wf.data_ptr(), cs.data_ptr(), bf.data_ptr(), stream_ptr(dev))
    _lib.check(rc, "fold_layernorm")
    return wf, cs, bf


def cast_weight(w: torch.Tensor, prec: str) -> torch.Tensor:
    """Matrix operand in the act dtype ([out,in] row-major, exactly as nn.Linear stores it)."""
    w = w.detach()
    if prec == "bf16":
        if w.dtype == torch.bfloat16:
            return w.contiguous()
        src = w.float().contiguous()
        out = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
        with torch.cuda.device(src.device):          # the module may live on a device that is not the current one
            rc = _lib.load().hsenet_cast_bf16(src.data_ptr(), out.data_ptr(), src.numel(), stream_ptr(src.device))
        _lib.check(rc, "cast_bf16")
        return out
    return w.float().contiguous()


def f32(p: torch.Tensor) -> torch.Tensor:
    return p.detach().float().contiguous()


_generation = 0


def _next_generation() -> int:
    global _generation
    _generation += 1
    return _generation


class WeightCache:
    """Derived (packed / down-cast) copies of a module's parameters, rebuilt when any parameter's storage or
    version counter changes (load_state_dict, optimizer step, .to()).

    ``generation`` is a process-wide monotonically increasing number bumped on every rebuild: anything that bakes the
    payload's device pointers in (captured CUDA graphs) validates against it -- never against ``id(payload)``, which
    CPython reuses once the previous payload is freed.

    Limitation: an in-place write through ``param.data`` (``p.data.copy_()``, EMA / clipping code, some checkpoint
    loaders) changes neither ``data_ptr`` nor ``_version``; call ``invalidate()`` (modules expose it as
    ``refresh_weights()``) after such an update."""

    def __init__(self):
        self.sig = None
        self.prec = None
        self.payload = None
        self.generation = 0

    def invalidate(self):
        self.sig = None
        self.payload = None

    def __deepcopy__(self, memo):
        return WeightCache()

    def __getstate__(self):
        return {}

    def __setstate__(self, state):
        self.__init__()

    def get(self, params, prec: str, builder):
        params = list(params)
        sig = _sig(params)
        if self.payload is None or sig != self.sig or prec != self.prec:
            self.payload = builder(prec)
            self.sig = sig
            self.prec = prec
            self.generation = _next_generation()
        return self.payload


_graph_kernels = 0


def note_graph_kernels(n: int) -> None:
    """Kernels launched on the device through CUDA-graph replays (the library's own counter only sees direct calls)."""
    global _graph_kernels
    _graph_kernels += n


def kernel_launch_count() -> int:
    """Total hsenet_b200 kernels launched by this process: direct launches + kernels inside replayed graphs."""
    return int(_lib.load().hsenet_launch_count()) + _graph_kernels


class GraphCache:
    """Captured CUDA graphs of a module's forward, keyed by the call shape; an entry is only valid for the weight-cache
    GENERATION and workspace address it was captured with (both are baked into the graph as raw pointers).  The entry
    also keeps a strong reference to the payload and the workspace, so the memory a live graph points at cannot be
    freed and re-used under it."""

    def __init__(self, max_entries: int = 8):
        self.entries = {}
        self.max_entries = max_entries

    def __deepcopy__(self, memo):
        return GraphCache(self.max_entries)

    def __getstate__(self):
        return {"max_entries": self.max_entries}

    def __setstate__(self, state):
        self.__init__(state.get("max_entries", 8))

    def get(self, key, generation, ws_ptr):
        ent = self.entries.get(key)
        if ent is None:
            return None
        if ent["_generation"] != generation or ent["_ws_ptr"] != ws_ptr:
            self.entries.pop(key, None)          # stale: weights were rebuilt or the workspace moved
            return None
        return ent

    def put(self, key, generation, ws_ptr, ent, keep=()):
        if len(self.entries) >= self.max_entries and key not in self.entries:
            self.entries.pop(next(iter(self.entries)))
        ent["_generation"] = generation
        ent["_ws_ptr"] = ws_ptr
        ent["_keep"] = tuple(keep)
        self.entries[key] = ent
        return ent

    def clear(self):
        self.entries.clear()


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()
