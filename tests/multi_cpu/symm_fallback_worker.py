"""Worker of tests/test_symmetric_fallback.py: drives SymmetricBuffer's handle-exchange protocol on CPU ranks
(gloo) with a fake driver whose calls can be made to fail on chosen ranks, and prints which heap every rank
ended up with.  Launched by torch.distributed.run."""
import json
import os
import sys
import tempfile

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from byzpy_b200.parallel.symmetric import SymmetricBuffer  # noqa: E402


class FakeDriver:
    """Mimics the VMM / multicast / IPC entry points of byzpy_b200._C.  Handles are small integers, exported
    file descriptors are real (temp files holding the handle id) so they really travel over SCM_RIGHTS."""

    def __init__(self, rank, fail):
        self.rank, self.fail = rank, fail           # fail: {call name: [ranks]}
        self.calls = []
        self.live_handles = set()
        self.live_maps = set()
        self._next = 1000 * (rank + 1)

    def _maybe_fail(self, name):
        self.calls.append(name)
        if self.rank in self.fail.get(name, ()):
            raise RuntimeError(f"injected failure in {name} on rank {self.rank}")

    def _new(self):
        self._next += 1
        return self._next

    def vmm_support(self, dev):
        return {"vmm": True, "posix_fd": True, "multicast": self.rank not in self.fail.get("no_mc_support", ()),
                "reason": ""}

    def vmm_granularity(self, dev, world, mc):
        self._maybe_fail("vmm_granularity")
        return 4096

    def vmm_alloc(self, nbytes, gran, dev):
        self._maybe_fail("vmm_alloc")
        h = self._new()
        self.live_handles.add(h)
        p = self._new()
        self.live_maps.add(p)
        return h, p

    def vmm_export_fd(self, h):
        self._maybe_fail("vmm_export_fd")
        f = tempfile.TemporaryFile()
        f.write(str(h).encode())
        f.flush()
        fd = os.dup(f.fileno())
        f.close()
        return fd

    def vmm_import_fd(self, fd, dev):
        self._maybe_fail("vmm_import_fd")
        # (descriptors received over SCM_RIGHTS share ONE file offset with every other receiver: read positionally)
        remote = int(os.pread(fd, 64, 0).decode())
        h = self._new()
        self.live_handles.add(h)
        self.imported = getattr(self, "imported", []) + [remote]
        return h

    def close_fd(self, fd):
        os.close(fd)

    def vmm_map(self, h, nbytes, gran, dev):
        self._maybe_fail("vmm_map_mc" if h == getattr(self, "mc", None) else "vmm_map")
        p = self._new()
        self.live_maps.add(p)
        return p

    def vmm_unmap(self, p, nbytes):
        self.live_maps.discard(p)

    def vmm_release(self, h):
        self.live_handles.discard(h)

    def mc_create(self, world, nbytes):
        self._maybe_fail("mc_create")
        h = self._new()
        self.live_handles.add(h)
        self.mc = h
        return h

    def mc_add_device(self, h, dev):
        self.mc = h
        self._maybe_fail("mc_add_device")

    def mc_bind(self, h, mem, nbytes):
        self._maybe_fail("mc_bind")

    def mc_unbind(self, h, dev, nbytes):
        self.calls.append("mc_unbind")

    # CUDA-IPC heap
    def raw_alloc(self, nbytes):
        p = self._new()
        self.live_maps.add(p)
        return p

    def raw_free(self, p):
        self.live_maps.discard(p)

    def ipc_export(self, p):
        return str(p).encode()

    def ipc_open(self, handle):
        p = self._new()
        self.live_maps.add(p)
        return p

    def ipc_close(self, p):
        self.live_maps.discard(p)


def main():
    fail = json.loads(sys.argv[1])
    fail = {k: list(v) for k, v in fail.items()}
    multicast = {"none": None, "true": True, "false": False}[sys.argv[2]]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    drv = FakeDriver(rank, fail)
    out = {"rank": rank}
    try:
        import warnings

        with warnings.catch_warnings(record=True) as caught:
            warnings.simplefilter("always")
            buf = SymmetricBuffer(10_000, torch.device("cpu"), kind="vmm", multicast=multicast, _ext=drv)
        out.update(kind=buf.kind, mc=bool(buf.mc_base), ptrs_ok=all(p != 0 for p in buf.ptrs),
                   distinct=len(set(buf.ptrs)) == world, warned=[str(w.message)[:60] for w in caught])
        dist.barrier()
        buf.close()
        out.update(leaked_maps=len(drv.live_maps), leaked_handles=len(drv.live_handles))
    except Exception as exc:       # every rank must fail together (no rank left waiting in a collective)
        out.update(error=type(exc).__name__ + ": " + str(exc)[:80])
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        print("RESULT " + json.dumps(gathered))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
