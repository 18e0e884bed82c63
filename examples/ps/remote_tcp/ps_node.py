"""Standalone TCP parameter server / honest worker / Byzantine worker (no actors, no scheduler):
the minimal multi-host deployment, one process per role, HMAC-authenticated frames.

Frame = 4-byte big-endian length | 32-byte HMAC-SHA256(secret, body) | body (pickle).  The secret
comes from ``BYZPY_HMAC_SECRET`` (same variable as the reference example); frames with a bad MAC are
dropped before unpickling.  Protocol (server <-> worker):

    worker -> {"type": "hello", "id": ...}                      server -> {"type": "init_model", "state_dict": {k: cpu}}
    server -> {"type": "round", "round": r}                     worker -> {"type": "gradient", "round": r, "vector": g}
    server -> {"type": "update", "round": r, "vector": agg}     (workers apply it with their own optimizer)
    server -> {"type": "done"}

A round closes when every connected worker has answered or after ``round_timeout`` seconds, and is skipped
when fewer gradients than the aggregator needs arrived; a worker whose connection drops is removed and the
training continues with the rest (counterpart of the reference's examples/ps/remote_tcp/ps_node.py).
Config keys: ``server``, ``workers`` (``id``, ``role``, optional ``leave_after``), ``rounds``, ``round_timeout``,
``lr``, ``eval_every``, ``aggregator``.

    python examples/ps/remote_tcp/ps_node.py server &
    for w in w0 w1 w2 w3; do python examples/ps/remote_tcp/ps_node.py worker --id $w & done; wait

The reference's spelling works too (``--worker-id`` = an id or a position in the list; ``--worker-type``,
``--rounds``, ``--data-root`` override the node list):

    python examples/ps/remote_tcp/ps_node.py --config nodes.yaml --role server
    python examples/ps/remote_tcp/ps_node.py --config nodes.yaml --role worker --worker-id 3 --worker-type byzantine
"""
from __future__ import annotations

import argparse
import asyncio
import hashlib
import hmac
import os
import pickle
import struct
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..")))

from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian, CoordinateWiseTrimmedMean  # noqa: E402
from byzpy_b200.aggregators.geometric_wise import MultiKrum  # noqa: E402
from byzpy_b200.attacks import EmpireAttack  # noqa: E402
from byzpy_b200.models import SmallCNN  # noqa: E402
from byzpy_b200.parallel.arena import flatten_grads, write_vector_to_grads_  # noqa: E402
from byzpy_b200.utils.data import batch_source, evaluate, mnist_like, shard_indices  # noqa: E402

SECRET = os.environ.get("BYZPY_HMAC_SECRET", "change-me").encode()


# ------------------------------------------------------------------------------ framing
async def send_frame(writer: asyncio.StreamWriter, obj) -> None:
    body = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    mac = hmac.new(SECRET, body, hashlib.sha256).digest()
    writer.write(struct.pack(">I", len(body)) + mac + body)
    await writer.drain()


MAX_FRAME = 256 << 20


async def recv_frame(reader: asyncio.StreamReader):
    (n,) = struct.unpack(">I", await reader.readexactly(4))
    if n > MAX_FRAME:
        raise PermissionError(f"frame of {n} bytes exceeds the {MAX_FRAME} byte limit")
    mac = await reader.readexactly(32)
    body = await reader.readexactly(n)
    if not hmac.compare_digest(mac, hmac.new(SECRET, body, hashlib.sha256).digest()):
        raise PermissionError("bad HMAC: frame dropped")
    return pickle.loads(body)


def load_config(path):
    import yaml

    with open(path) as f:
        return yaml.safe_load(f)


def build_aggregator(spec):
    name = spec.get("name", "median")
    if name == "median":
        return CoordinateWiseMedian()
    if name == "trimmed_mean":
        return CoordinateWiseTrimmedMean(f=int(spec.get("f", 1)))
    if name == "multi_krum":
        return MultiKrum(f=int(spec.get("f", 1)), q=int(spec.get("q", 2)))
    raise ValueError(f"unknown aggregator {name!r}")


# ------------------------------------------------------------------------------ server
class ParameterServerNode:
    """The server role: accepts workers, drives rounds, aggregates, broadcasts the update.

    One receive task per worker connection feeds a per-round table, so a slow or dead worker never blocks
    the others: a round closes when every CONNECTED worker has answered or ``round_timeout`` expires; a
    worker whose connection drops is removed from the table (the round proceeds with the rest); a round
    with too few gradients for the aggregator is skipped (the reference skips on a 60 s timeout,
    reference examples/ps/remote_tcp/ps_node.py:389-394)."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.agg = build_aggregator(cfg.get("aggregator", {}))
        self.expected = {str(w["id"]) for w in cfg["workers"]}
        self.lr = float(cfg.get("lr", 0.05))
        self.rounds = int(cfg.get("rounds", 5))
        self.timeout = float(cfg.get("round_timeout", 60))
        self.eval_every = int(cfg.get("eval_every", 1))
        torch.manual_seed(0)
        self.model = SmallCNN()
        self.opt = torch.optim.SGD(self.model.parameters(), lr=self.lr)
        self.writers = {}                  # worker id -> StreamWriter
        self.tasks = {}                    # worker id -> receive task
        self.round = 0
        self.inbox = {}                    # worker id -> gradient of the current round
        self.changed = asyncio.Event()     # a gradient arrived or a worker left
        self.all_in = asyncio.Event()
        self.rejected = 0
        self._srv = None

    async def start(self):
        host, port = self.cfg["server"]["host"], int(self.cfg["server"]["port"])
        self._srv = await asyncio.start_server(self._on_conn, host, port)
        print(f"parameter server on {host}:{port}, waiting for {len(self.expected)} workers", flush=True)

    async def _on_conn(self, reader, writer):
        try:
            hello = await recv_frame(reader)
            wid = str(hello["id"])
            if hello.get("type") != "hello" or wid not in self.expected:
                raise PermissionError(f"unexpected worker {wid!r}")
        except Exception as exc:  # noqa: BLE001   (bad MAC, unknown id, garbage)
            self.rejected += 1
            print("rejected connection:", exc, flush=True)
            writer.close()
            return
        self.writers[wid] = writer
        await send_frame(writer, {"type": "init_model", "round": self.round,
                                  "state_dict": {k: v.cpu() for k, v in self.model.state_dict().items()}})
        self.tasks[wid] = asyncio.ensure_future(self._receive(wid, reader))
        if self.expected <= set(self.writers):
            self.all_in.set()

    async def _receive(self, wid, reader):
        try:
            while True:
                msg = await recv_frame(reader)
                if msg.get("type") == "gradient" and msg.get("round") == self.round:
                    self.inbox[wid] = msg["vector"]
                    self.changed.set()
        except (asyncio.IncompleteReadError, ConnectionError, PermissionError, OSError) as exc:
            print(f"worker {wid} left ({type(exc).__name__})", flush=True)
        finally:
            self.writers.pop(wid, None)
            self.changed.set()

    async def _broadcast(self, msg):
        for wid, w in list(self.writers.items()):
            try:
                await send_frame(w, msg)
            except (ConnectionError, OSError):
                self.writers.pop(wid, None)

    async def _collect(self):
        loop = asyncio.get_running_loop()
        deadline = loop.time() + self.timeout
        while loop.time() < deadline and self.writers and not set(self.writers) <= set(self.inbox):
            self.changed.clear()
            try:
                await asyncio.wait_for(self.changed.wait(), timeout=max(0.01, deadline - loop.time()))
            except asyncio.TimeoutError:
                break
        return [self.inbox[w] for w in sorted(self.inbox)]          # deterministic row order

    async def run_training(self):
        await self.all_in.wait()
        xt, yt = mnist_like(2000, train=False, root=self.cfg.get("data_root", "./data"))
        for r in range(1, self.rounds + 1):
            self.round, self.inbox = r, {}
            await self._broadcast({"type": "round", "round": r})
            grads = await self._collect()
            try:
                vec = self.agg.aggregate(grads)
            except ValueError as exc:       # not enough gradients for this aggregator
                print(f"[round {r}] skipped ({len(grads)} gradients: {exc})", flush=True)
                continue
            write_vector_to_grads_(self.model, vec)
            self.opt.step()
            await self._broadcast({"type": "update", "round": r, "vector": vec})
            if r % self.eval_every == 0 or r == self.rounds:
                loss, acc = evaluate(self.model, xt, yt, torch.device("cpu"))
                print(f"[round {r}] {len(grads)} gradients  test loss={loss:.4f} acc={acc:.4f}", flush=True)
        await self._broadcast({"type": "done"})

    async def shutdown(self):
        for w in list(self.writers.values()):
            w.close()
        for t in self.tasks.values():
            t.cancel()
        if self._srv is not None:
            self._srv.close()
            try:        # (3.12 waits here for every accepted connection to be gone; do not hang on a stuck peer)
                await asyncio.wait_for(self._srv.wait_closed(), timeout=2.0)
            except asyncio.TimeoutError:
                pass


async def run_server(cfg):
    node = ParameterServerNode(cfg)
    await node.start()
    try:
        await node.run_training()
    finally:
        await node.shutdown()


# ------------------------------------------------------------------------------ workers
def resolve_worker(cfg, wid: str) -> dict:
    """The worker entry named ``wid``; a bare number that is nobody's id is a position in the list
    (``--worker-id 0`` as in the reference's command lines)."""
    for w in cfg["workers"]:
        if str(w["id"]) == str(wid):
            return w
    if str(wid).isdigit() and int(wid) < len(cfg["workers"]):
        return cfg["workers"][int(wid)]
    raise SystemExit(f"no worker {wid!r} in the node list ({[str(w['id']) for w in cfg['workers']]})")


async def run_worker(cfg, wid):
    entry = resolve_worker(cfg, wid)
    wid = str(entry["id"])
    honest_ids = [str(w["id"]) for w in cfg["workers"] if w.get("role", "honest") == "honest"]
    for attempt in range(50):
        try:
            reader, writer = await asyncio.open_connection(cfg["server"]["host"], int(cfg["server"]["port"]))
            break
        except OSError:
            await asyncio.sleep(0.2)
    else:
        raise ConnectionError("parameter server not reachable")
    await send_frame(writer, {"type": "hello", "id": wid})
    init = await recv_frame(reader)
    model = SmallCNN()
    model.load_state_dict(init["state_dict"], strict=True)
    opt = torch.optim.SGD(model.parameters(), lr=float(cfg.get("lr", 0.05)))
    lossf = torch.nn.CrossEntropyLoss()
    byz = entry.get("role", "honest") != "honest"
    leave_after = entry.get("leave_after")          # drop the connection after this many rounds (failure rehearsal)
    if byz:
        me = 0
        attack = EmpireAttack(scale=-1.0)
    else:
        me = honest_ids.index(wid)
    x, y = mnist_like(6000, root=cfg.get("data_root", "./data"))
    idx = torch.as_tensor(shard_indices(6000, max(1, len(honest_ids)))[me])
    nxt = batch_source(x[idx], y[idx], 64, seed=me)
    while True:
        msg = await recv_frame(reader)
        if msg["type"] == "done":
            break
        if msg["type"] == "round":
            if leave_after is not None and msg["round"] > int(leave_after):
                print(f"[{wid}] leaving before round {msg['round']}", flush=True)
                break
            xb, yb = nxt()
            model.zero_grad(set_to_none=True)
            lossf(model(xb), yb).backward()
            g = flatten_grads(model)
            if byz:     # a non-omniscient Empire: scale its own honest gradient
                g = torch.as_tensor(attack.apply(honest_grads=[g]))
            await send_frame(writer, {"type": "gradient", "round": msg["round"], "vector": g})
        elif msg["type"] == "update":
            write_vector_to_grads_(model, msg["vector"])
            opt.step()
    writer.close()
    print(f"[{wid}] finished", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("role_pos", nargs="?", choices=["server", "worker"], default=None, metavar="role")
    ap.add_argument("--role", choices=["server", "worker"], default=None)
    ap.add_argument("--id", "--worker-id", dest="id", default="w0",
                    help="worker id from the node list, or its position in it")
    ap.add_argument("--worker-type", choices=["honest", "byzantine"], default=None,
                    help="override this worker's role from the node list")
    ap.add_argument("--rounds", type=int, default=None, help="override the node list's `rounds`")
    ap.add_argument("--data-root", default=None, help="MNIST directory (synthetic stand-in when absent)")
    ap.add_argument("--config", default=os.path.join(os.path.dirname(__file__), "nodes_example.yaml"))
    a = ap.parse_args()
    role = a.role or a.role_pos
    if role is None:
        ap.error("say which node this is: --role server | --role worker")
    cfg = load_config(a.config)
    if a.rounds is not None:
        cfg["rounds"] = a.rounds
    if a.data_root is not None:
        cfg["data_root"] = a.data_root
    if role == "worker" and a.worker_type is not None:
        resolve_worker(cfg, a.id)["role"] = a.worker_type
    asyncio.run(run_server(cfg) if role == "server" else run_worker(cfg, a.id))
