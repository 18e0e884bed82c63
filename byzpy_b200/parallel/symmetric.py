"""Symmetric device memory (one process per GPU).

Every rank allocates an identically sized raw arena and maps every peer's arena into its own
address space; the fused kernels then address peer HBM directly (P2P ``ld.global`` /
``st.global`` over NVLink 5 / NVSwitch).  Two heaps:

* ``"vmm"`` (default where supported) -- CUDA virtual-memory-management allocations
  (``cuMemCreate``) exported as POSIX file descriptors, exchanged between the ranks over Unix
  sockets (``SCM_RIGHTS``), and bound to one **NVLS multicast object** for the whole team
  (``cuMulticastCreate`` / ``AddDevice`` / ``BindMem``).  ``mc_ptr()`` is the multicast alias: a
  ``multimem.st`` there is replicated by the switch into every rank's copy, a
  ``multimem.ld_reduce`` sums the copies in the switch (``csrc/vmm.cpp``).
* ``"ipc"`` -- ``cudaMalloc`` + ``cudaIpc*`` handles (``csrc/runtime.cpp``); no multicast.  The
  fallback when the driver lacks VMM / fd export, and selectable with ``BYZPY_SYMM=ipc``.

``torch.distributed`` (NCCL or Gloo) is used ONLY here, at start-up, to exchange handles / socket
paths.

B200-native replacement for the reference's host shared-memory store and pickled tensor
transport (reference engine/storage/shared_store.py:21-54, engine/actor/ipc.py:20-56,
engine/actor/transports/ucx.py:225-270).
"""
from __future__ import annotations

import os
import socket
import tempfile
import threading
import time
import uuid
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


# --------------------------------------------------------------------- fd exchange
class _FdServer:
    """Serves this rank's file descriptors to its peers over a Unix socket (``SCM_RIGHTS``).
    A request is one byte: the index of the wanted descriptor."""

    def __init__(self, fds: List[int]):
        self.fds = list(fds)
        # Linux abstract namespace (leading NUL): no file, so neither a full / read-only / very long temp
        # directory nor a stale path can make the exchange -- and with it the VMM heap -- fail
        self.path = f"\0byzpy_b200_fd_{os.getpid()}_{uuid.uuid4().hex}"
        self._sock = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        try:
            self._sock.bind(self.path)
        except OSError:                                    # (a kernel without the abstract namespace)
            self.path = os.path.join(tempfile.gettempdir(), f"byzpy_b200_fd_{uuid.uuid4().hex}.sock")
            self._sock.bind(self.path)
        self._sock.listen(64)
        self._stop = False
        self._thr = threading.Thread(target=self._serve, daemon=True)
        self._thr.start()

    def _serve(self) -> None:
        while not self._stop:
            try:
                conn, _ = self._sock.accept()
            except OSError:
                return
            with conn:
                try:
                    idx = conn.recv(1)
                    if idx:
                        socket.send_fds(conn, [b"f"], [self.fds[idx[0]]])
                except OSError:
                    pass

    def close(self) -> None:
        self._stop = True
        try:
            self._sock.close()
        finally:
            if not self.path.startswith("\0"):
                try:
                    os.unlink(self.path)
                except OSError:
                    pass


def _fetch_fd(path: str, index: int = 0, attempts: int = 5) -> int:
    """Ask the peer serving ``path`` for its descriptor ``index``.  A refused / interrupted exchange (a loaded box:
    full accept backlog, EINTR, a descriptor table momentarily full) is retried a few times before it counts as
    a failure -- one failed fetch moves the WHOLE team off the VMM heap."""
    last: Optional[BaseException] = None
    for k in range(attempts):
        try:
            with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as s:
                s.connect(path)
                s.sendall(bytes([index]))
                _, fds, _, _ = socket.recv_fds(s, 1, 1)
                if fds:
                    return fds[0]
                last = RuntimeError(f"no file descriptor received from {path!r}")
        except OSError as exc:
            last = exc
        time.sleep(0.05 * (k + 1))
    raise RuntimeError(f"could not fetch descriptor {index} from {path!r}: {last!r}")


def heap_kind(device: torch.device, world: int) -> str:
    """``"vmm"`` when the driver supports VMM allocations with POSIX-fd export, else ``"ipc"``."""
    forced = os.environ.get("BYZPY_SYMM", "").lower()
    if forced in ("ipc", "vmm"):
        return forced
    ext = ops.require_ext()
    sup = ext.vmm_support(device.index if device.index is not None else torch.cuda.current_device())
    return "vmm" if (sup["vmm"] and sup["posix_fd"]) else "ipc"


class SymmetricBuffer:
    """A raw device allocation mapped on every rank of ``group``.

    ``ptrs[r]`` is the address at which rank ``r``'s copy is visible in THIS process
    (``ptrs[rank]`` is the local allocation); ``mc_base`` the address of the NVLS multicast alias
    (0 without one).  ``local`` is a uint8 torch tensor over the local copy; use :meth:`view` for
    typed views.
    """

    def __init__(self, nbytes: int, device: torch.device, group=None, *, kind: Optional[str] = None,
                 multicast: Optional[bool] = None, _ext=None):
        ext = _ext if _ext is not None else ops.require_ext()     # (_ext: a fake driver, CPU protocol tests)
        self._ext = ext
        self.device = device
        self.group = group
        self.rank = dist.get_rank(group) if _dist_on() else 0
        self.world = dist.get_world_size(group) if _dist_on() else 1
        self._dev_index = (device.index if device.index is not None
                           else (torch.cuda.current_device() if device.type == "cuda" else 0))
        self.kind = kind or heap_kind(device, self.world)
        self.mc_base = 0
        self._closed = False
        self._opened: List[int] = []
        self.ptrs: List[int] = [0] * self.world
        if self.world > 1:      # every rank must take the same path
            kinds: List[Optional[str]] = [None] * self.world
            dist.all_gather_object(kinds, self.kind, group=group)
            if any(k != "vmm" for k in kinds):
                self.kind = "ipc"
        if self.kind == "vmm":
            self._init_vmm(nbytes, multicast)
        else:
            if multicast:
                raise RuntimeError("NVLS multicast needs the VMM symmetric heap")
            self._init_ipc(nbytes)
        self.local = (tensor_from_ptr(self._ptr, self.nbytes, device, owner=self) if device.type == "cuda"
                      else torch.empty(0, dtype=torch.uint8))

    # ----------------------------------------------------------------- CUDA IPC heap
    def _init_ipc(self, nbytes: int) -> None:
        ext = self._ext
        self.nbytes = int((nbytes + 255) // 256 * 256)
        with _device_ctx(self.device):
            self._ptr = ext.raw_alloc(self.nbytes)
        self.ptrs[self.rank] = self._ptr
        if self.world > 1:
            handle = ext.ipc_export(self._ptr)
            handles: List[Optional[bytes]] = [None] * self.world
            dist.all_gather_object(handles, handle, group=self.group)
            with _device_ctx(self.device):
                for r, h in enumerate(handles):
                    if r == self.rank:
                        continue
                    p = ext.ipc_open(h)
                    self.ptrs[r] = p
                    self._opened.append(p)

    # ----------------------------------------------------------------- VMM heap (+ multicast)
    def _agree(self, ok: bool) -> bool:
        """True when ``ok`` holds on EVERY rank.  Each set-up stage ends with exactly one of these, whether
        or not it failed locally, so the ranks never diverge in their sequence of collectives and the
        whole team takes the same fallback."""
        if self.world == 1:
            return bool(ok)
        flags: List[Optional[bool]] = [None] * self.world
        dist.all_gather_object(flags, bool(ok), group=self.group)
        return all(flags)

    def _init_vmm(self, nbytes: int, multicast: Optional[bool]) -> None:
        """VMM heap with the fallbacks agreed team-wide, stage by stage: no multicast support, or a
        multicast step failing on any rank (no NVLS on this fabric, fabric manager down) -> unicast VMM
        heap, deliveries by peer stores; allocation / fd export / import / mapping failing anywhere ->
        the CUDA-IPC heap.  ``multicast=True`` turns the multicast fallbacks into errors."""
        import warnings

        ext, dev = self._ext, self._dev_index
        sup = ext.vmm_support(dev)
        if multicast and self.world > 1 and not sup["multicast"]:
            raise RuntimeError("multicast=True but this device / driver reports no NVLS multicast support")
        want_mc = self._agree(self.world > 1 and multicast is not False and bool(sup["multicast"])) \
            if self.world > 1 else False
        self._peer_handles: List[int] = []
        self._mc_handle = 0
        self._mc_bound = False
        self._handle = 0
        self._ptr = 0
        fds: List[int] = []
        err: Optional[BaseException] = None
        # ---- stage A: local allocation, export, (rank 0) the multicast object
        mc_local_ok = True
        try:
            gran = int(ext.vmm_granularity(dev, self.world, want_mc))
            self._gran = gran
            self.nbytes = int((nbytes + gran - 1) // gran * gran)
            self._handle, self._ptr = ext.vmm_alloc(self.nbytes, gran, dev)
            self.ptrs[self.rank] = self._ptr
            if self.world > 1:
                fds.append(ext.vmm_export_fd(self._handle))
                if want_mc and self.rank == 0:
                    try:
                        self._mc_handle = ext.mc_create(self.world, self.nbytes)
                        fds.append(ext.vmm_export_fd(self._mc_handle))
                    except Exception as exc:            # no multicast object: carry on unicast
                        mc_local_ok = False
                        err = exc
        except Exception as exc:
            err = exc
            mc_local_ok = False
        if self.world == 1:
            if self._ptr == 0:
                self._vmm_teardown(fds)
                warnings.warn(f"VMM allocation failed ({err!r}); using the CUDA-IPC heap")
                self._fallback_ipc(nbytes, multicast)
            return
        if not self._agree(self._ptr != 0 and len(fds) >= 1):
            self._vmm_teardown(fds)
            warnings.warn(f"VMM symmetric heap unavailable on some rank ({err!r} here); using the CUDA-IPC heap")
            self._fallback_ipc(nbytes, multicast)
            return
        if want_mc and not self._agree(mc_local_ok):
            if multicast:
                self._vmm_teardown(fds)
                raise RuntimeError(f"multicast=True but the multicast object could not be created: {err!r}")
            self._drop_multicast()
            want_mc = False
        # ---- stage B: exchange the allocation handles as file descriptors, map every peer
        server = None
        try:
            try:
                server = _FdServer(fds)
                path: Optional[str] = server.path
            except Exception as exc:
                err, path = exc, None
            paths: List[Optional[str]] = [None] * self.world
            dist.all_gather_object(paths, path, group=self.group)
            ok = all(p is not None for p in paths)
            if ok:
                try:
                    for r in range(self.world):
                        if r == self.rank:
                            continue
                        fd = _fetch_fd(paths[r], 0)
                        try:
                            h = ext.vmm_import_fd(fd, dev)
                        finally:
                            ext.close_fd(fd)
                        self._peer_handles.append(h)
                        p = ext.vmm_map(h, self.nbytes, self._gran, dev)
                        self.ptrs[r] = p
                        self._opened.append(p)
                    if want_mc and self.rank != 0:
                        fd = _fetch_fd(paths[0], 1)
                        try:
                            self._mc_handle = ext.vmm_import_fd(fd, dev)
                        finally:
                            ext.close_fd(fd)
                except Exception as exc:
                    err, ok = exc, False
            if not self._agree(ok):     # (also: nobody tears its server down while peers still fetch)
                self._vmm_teardown(fds)
                fds = []
                warnings.warn(f"VMM handle exchange failed on some rank ({err!r} here); using the CUDA-IPC heap")
                self._fallback_ipc(nbytes, multicast)
                return
            # ---- stage C: bind everybody's allocation to the multicast object, map its alias
            if want_mc:
                steps = (lambda: ext.mc_add_device(self._mc_handle, dev),      # every device joins the team ...
                         lambda: self._mc_bind(),                              # ... before anyone binds
                         lambda: setattr(self, "mc_base", ext.vmm_map(self._mc_handle, self.nbytes, self._gran, dev)))
                for step in steps:
                    good = True
                    try:
                        step()
                    except Exception as exc:
                        err, good = exc, False
                    if not self._agree(good):
                        if multicast:
                            self._vmm_teardown(fds)
                            fds = []
                            raise RuntimeError(f"multicast=True but the NVLS set-up failed: {err!r}")
                        warnings.warn(f"NVLS multicast unavailable ({err!r} here); deliveries use peer stores")
                        self._drop_multicast()
                        break
        finally:
            if server is not None:
                server.close()
            for fd in fds:
                try:
                    ext.close_fd(fd)
                except Exception:
                    pass

    def _mc_bind(self) -> None:
        self._ext.mc_bind(self._mc_handle, self._handle, self.nbytes)
        self._mc_bound = True

    def _drop_multicast(self) -> None:
        ext = self._ext
        try:
            if self.mc_base:
                ext.vmm_unmap(self.mc_base, self.nbytes)
            if self._mc_handle:
                if self._mc_bound:
                    ext.mc_unbind(self._mc_handle, self._dev_index, self.nbytes)
                ext.vmm_release(self._mc_handle)
        except Exception:
            pass
        self.mc_base = 0
        self._mc_handle = 0
        self._mc_bound = False

    def _vmm_teardown(self, fds: List[int]) -> None:
        """Undo a partially built VMM heap (best effort) before falling back."""
        ext = self._ext
        self._drop_multicast()
        for p in self._opened:
            try:
                ext.vmm_unmap(p, self.nbytes)
            except Exception:
                pass
        for h in self._peer_handles:
            try:
                ext.vmm_release(h)
            except Exception:
                pass
        self._opened, self._peer_handles = [], []
        try:
            if self._ptr:
                ext.vmm_unmap(self._ptr, self.nbytes)
            if self._handle:
                ext.vmm_release(self._handle)
        except Exception:
            pass
        self._ptr = self._handle = 0
        for fd in fds:
            try:
                ext.close_fd(fd)
            except Exception:
                pass
        self.ptrs = [0] * self.world

    def _fallback_ipc(self, nbytes: int, multicast: Optional[bool]) -> None:
        if multicast:
            raise RuntimeError("NVLS multicast needs the VMM symmetric heap")
        self.kind = "ipc"
        self._init_ipc(nbytes)

    # ----------------------------------------------------------------- views
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
        return (self.mc_base + offset_bytes) if self.mc_base else 0

    def close(self) -> None:
        if self._closed:
            return
        self._closed = True
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)
        ext = self._ext
        with _device_ctx(self.device):
            if self.world > 1 and _dist_on():
                try:
                    dist.barrier(group=self.group)      # peers have stopped touching my memory
                except Exception:
                    pass
            if self.kind == "vmm":
                try:
                    self._drop_multicast()
                    for p in self._opened:
                        ext.vmm_unmap(p, self.nbytes)
                    for h in self._peer_handles:
                        ext.vmm_release(h)
                    ext.vmm_unmap(self._ptr, self.nbytes)
                    ext.vmm_release(self._handle)
                except Exception:
                    pass
                return
            for p in self._opened:
                try:
                    ext.ipc_close(p)
                except Exception:
                    pass
            if self.world > 1 and _dist_on():
                try:
                    dist.barrier(group=self.group)
                except Exception:
                    pass
            try:
                ext.raw_free(self._ptr)
            except Exception:
                pass


def _device_ctx(device: torch.device):
    import contextlib

    return torch.cuda.device(device) if device.type == "cuda" else contextlib.nullcontext()


def _dist_on() -> bool:
    return dist.is_available() and dist.is_initialized()


__all__ = ["SymmetricBuffer", "tensor_from_ptr", "heap_kind"]
