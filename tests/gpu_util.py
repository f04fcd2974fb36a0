"""Helpers for the -m gpu parity tests: numpy <-> device tensors and raw C-ABI stage calls."""
import ctypes

import numpy as np
import torch

from mhada_style_transfer_b200 import _lib

DEV = "cuda:0"


def tdt(code):
    return torch.bfloat16 if code == _lib.BF16 else torch.float32


def to_tokens(x: np.ndarray, code: int) -> torch.Tensor:
    """(B,C,h,w) numpy -> contiguous [B, h*w, C] device tensor of the path dtype."""
    b, c = x.shape[:2]
    t = torch.from_numpy(np.ascontiguousarray(x.reshape(b, c, -1).transpose(0, 2, 1))).to(torch.float32)
    return t.to(DEV).to(tdt(code)).contiguous()


def from_tokens(t: torch.Tensor, hw) -> np.ndarray:
    """[B, N, C] device tensor -> (B,C,h,w) float64 numpy."""
    b, n, c = t.shape
    return t.float().cpu().numpy().astype(np.float64).transpose(0, 2, 1).reshape(b, c, hw[0], hw[1])


def f32(x: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(DEV)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ws(nbytes: int) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=DEV)


def stats(x_tok: torch.Tensor, code: int):
    L = _lib.lib()
    B, N, C = x_tok.shape
    mean = torch.empty(B, C, dtype=torch.float32, device=DEV)
    rstd = torch.empty_like(mean)
    w = ws(L.mhada_in_stats_workspace(B, N, C))
    _lib.check("mhada_in_stats", L.mhada_in_stats(ptr(x_tok), code, B, N, C, C, ptr(mean), ptr(rstd), ptr(w),
                                                  w.numel(), stream()))
    return mean, rstd


def pack_fgh(sd: dict, H: int, prefix: str = ""):
    w = np.stack([np.stack([sd[f"{prefix}{n}.{i}.weight"].reshape(sd[f"{prefix}{n}.{i}.weight"].shape[0], -1)
                            for i in range(H)]) for n in ("f_list", "g_list", "h_list")])
    b = np.stack([np.stack([sd[f"{prefix}{n}.{i}.bias"] for i in range(H)]) for n in ("f_list", "g_list", "h_list")])
    return f32(w), f32(b)


def bf16_round(x: np.ndarray) -> np.ndarray:
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(torch.bfloat16).float().numpy().astype(np.float64)
