"""Peer-mapped device arenas for the multi-GPU sweep (one process per GPU).

Every rank allocates one buffer through the C ABI (gpa_peer_alloc: cudaMalloc + CUDA IPC handle),
the 64-byte handles are exchanged with one torch.distributed all-gather, and every rank maps the
buffers of its peers (gpa_peer_open).  After that the kernels of libgpa_b200.so address peer HBM
directly over NVLink; torch.distributed is not on the data path any more.  PyTorch only wraps the
LOCAL buffer as tensors (zero-copy, through __cuda_array_interface__).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _lib

__all__ = ["PeerArena", "MAX_PEERS"]

MAX_PEERS = 8
_HANDLE = 64


class _Blob:
    """Minimal __cuda_array_interface__ holder: lets torch adopt a raw device pointer without a copy."""

    def __init__(self, ptr, nbytes, owner):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2, "strides": None}
        self._owner = owner


class PeerArena:
    """`nbytes` of zero-initialised device memory on every rank of `group`, each mapped into all ranks.

    ptrs[r] is the address of rank r's buffer in THIS process (ptrs[rank] is the local allocation).
    Collective: every rank of the group must construct it, with the same size, in the same order."""

    def __init__(self, nbytes, group=None, device=None):
        self.lib = _lib.load()
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world > MAX_PEERS:
            raise ValueError(f"at most {MAX_PEERS} peers (one NVSwitch box), got {self.world}")
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.nbytes = int(nbytes)
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * _HANDLE)()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.gpa_peer_alloc(self.nbytes, ctypes.byref(ptr), handle))
        self._local = ptr.value
        self.ptrs = [None] * self.world
        self.ptrs[self.rank] = self._local
        self._opened = []
        if self.world > 1:
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.device)
            every = torch.empty(self.world * _HANDLE, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(every, mine, group=group)
            every = every.cpu().numpy().reshape(self.world, _HANDLE)
            for r in range(self.world):
                if r == self.rank:
                    continue
                h = (ctypes.c_ubyte * _HANDLE)(*every[r].tolist())
                p = ctypes.c_void_p()
                with torch.cuda.device(self.device):
                    _lib.check(self.lib.gpa_peer_open(h, ctypes.byref(p)))
                self.ptrs[r] = p.value
                self._opened.append(p.value)
            dist.barrier(group=group)      # nobody signals into an arena its owner has not zeroed and published yet
        self._bytes = torch.as_tensor(_Blob(self._local, self.nbytes, self), device=self.device)

    def tensor(self, offset, shape, dtype):
        """Zero-copy torch view of the LOCAL buffer at byte `offset`."""
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        if offset % 16 or offset + n > self.nbytes:
            raise ValueError("arena view out of range or misaligned")
        return self._bytes[offset:offset + n].view(dtype).view(*shape)

    def addr(self, r, offset=0):
        """Address of byte `offset` of rank r's buffer, valid in this process."""
        return self.ptrs[r] + int(offset)

    def close(self):
        if self._local is None:
            return
        torch.cuda.synchronize(self.device)
        if self.world > 1 and dist.is_initialized():
            dist.barrier(group=self.group)      # no peer is still writing into a buffer about to be unmapped
        for p in self._opened:
            self.lib.gpa_peer_close(ctypes.c_void_p(p))
        self._opened = []
        self._bytes = None
        self.lib.gpa_peer_free(ctypes.c_void_p(self._local))
        self._local = None

    def __del__(self):   # best effort; close() is the collective, orderly way
        try:
            if self._local is not None and self.world == 1:
                self.close()
        except Exception:
            pass
