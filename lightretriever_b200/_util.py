"""Small helpers shared by the host-side wrappers: torch is the tensor container, nothing more."""
from __future__ import annotations

import torch

from . import _C

_DTYPE_TAG = {torch.float32: _C.LR_F32, torch.bfloat16: _C.LR_BF16}


def dtype_tag(dt: torch.dtype) -> int:
    try:
        return _DTYPE_TAG[dt]
    except KeyError:
        raise ValueError(f"unsupported dtype {dt}: the B200 kernels take torch.float32 or torch.bfloat16") from None


def require_cuda(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t)}")
    if t.device.type != "cuda":
        raise RuntimeError(
            f"{name} is on {t.device}: lightretriever_b200 has no CPU path (move it to a B200 with .cuda())")
    return t


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


class Workspace:
    """Grow-only scratch buffers, one per (device, stream, host thread): the C ABI is re-entrant per (device, stream) and
    leaves the workspace to the caller, so two streams — or two host threads — must never share one.  Buffers are 256-byte
    aligned by the caching allocator; a buffer that grows is released in stream order (it was allocated on, and is only
    used by kernels of, the stream it is keyed by)."""

    def __init__(self) -> None:
        self._buf: dict[tuple, torch.Tensor] = {}

    def get(self, nbytes: int, device: torch.device) -> torch.Tensor:
        import threading

        device = torch.device(device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        key = (device, torch.cuda.current_stream(device).cuda_stream, threading.get_ident())
        buf = self._buf.get(key)
        if buf is None or buf.numel() < nbytes:
            self._buf.pop(key, None)
            with torch.cuda.device(device):
                buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._buf[key] = buf
        return buf

    def clear(self) -> None:
        self._buf.clear()
