"""Symmetric device memory over CUDA IPC (one process per GPU).

Every rank allocates identically sized raw arenas with ``cudaMalloc`` (through
the native runtime in ``byzpy_b200._C``), exports them as 64-byte IPC handles,
and maps every peer's arena into its own address space.  ``torch.distributed``
(NCCL or Gloo) is used ONLY here, to all-gather the handles at start-up; the
training hot path then addresses peer HBM directly from inside the fused
kernels (P2P ``ld.global``/``st.global`` over NVLink 5 / NVSwitch).

B200-native replacement for the reference's host shared-memory store and
pickled tensor transport (reference engine/storage/shared_store.py:21-54,
engine/actor/ipc.py:20-56, engine/actor/transports/ucx.py:225-270).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from .. import ops


class _RawView:
    """Expose a raw device pointer to torch through ``__cuda_array_interface__``."""

    def __init__(self, ptr: int, nbytes: int, owner):
        self._owner = owner  # keep the allocation alive
        self.__cuda_array_interface__ = {
            "shape": (nbytes,),
            "typestr": "|u1",
            "data": (ptr, False),
            "version": 2,
            "strides": None,
        }


def tensor_from_ptr(ptr: int, nbytes: int, device: torch.device, owner=None) -> torch.Tensor:
    """uint8 tensor aliasing ``nbytes`` of raw device memory at ``ptr`` (zero copy)."""
    return torch.as_tensor(_RawView(ptr, nbytes, owner), device=device)


class SymmetricBuffer:
    """A raw device allocation mapped on every rank of ``group``.

    ``ptrs[r]`` is the address at which rank ``r``'s copy is visible in THIS
    process (``ptrs[rank]`` is the local allocation).  ``local`` is a uint8
    torch tensor over the local copy; use :meth:`view` for typed views.
    """

    def __init__(self, nbytes: int, device: torch.device, group=None):
        ext = ops.require_ext()
        self._ext = ext
        self.nbytes = int((nbytes + 255) // 256 * 256)
        self.device = device
        self.group = group
        self.rank = dist.get_rank(group) if _dist_on() else 0
        self.world = dist.get_world_size(group) if _dist_on() else 1
        with torch.cuda.device(device):
            self._ptr = ext.raw_alloc(self.nbytes)
        self.ptrs: List[int] = [0] * self.world
        self.ptrs[self.rank] = self._ptr
        self._opened: List[int] = []
        if self.world > 1:
            handle = ext.ipc_export(self._ptr)
            handles: List[Optional[bytes]] = [None] * self.world
            dist.all_gather_object(handles, handle, group=group)
            with torch.cuda.device(device):
                for r, h in enumerate(handles):
                    if r == self.rank:
                        continue
                    p = ext.ipc_open(h)
                    self.ptrs[r] = p
                    self._opened.append(p)
        self.local = tensor_from_ptr(self._ptr, self.nbytes, device, owner=self)
        self._closed = False

    def view(self, dtype: torch.dtype, numel: Optional[int] = None, offset_bytes: int = 0) -> torch.Tensor:
        itemsize = torch.empty((), dtype=dtype).element_size()
        avail = (self.nbytes - offset_bytes) // itemsize
        numel = avail if numel is None else numel
        if numel > avail:
            raise ValueError("view exceeds symmetric buffer")
        return self.local[offset_bytes: offset_bytes + numel * itemsize].view(dtype)

    def peer_ptr(self, rank: int, offset_bytes: int = 0) -> int:
        return self.ptrs[rank] + offset_bytes

    def mc_ptr(self, offset_bytes: int = 0) -> int:
        """Address of the NVLS multicast alias of the buffer (a ``multimem.st`` there lands in every
        rank's copy), or 0 when the heap has no multicast mapping."""
        return (self.mc_base + offset_bytes) if getattr(self, "mc_base", 0) else 0

    def close(self) -> None:
        if self._closed:
            return
        self._closed = True
        torch.cuda.synchronize(self.device)
        with torch.cuda.device(self.device):
            for p in self._opened:
                try:
                    self._ext.ipc_close(p)
                except Exception:
                    pass
            if self.world > 1 and _dist_on():
                try:
                    dist.barrier(group=self.group)
                except Exception:
                    pass
            try:
                self._ext.raw_free(self._ptr)
            except Exception:
                pass


def _dist_on() -> bool:
    return dist.is_available() and dist.is_initialized()


__all__ = ["SymmetricBuffer", "tensor_from_ptr"]
