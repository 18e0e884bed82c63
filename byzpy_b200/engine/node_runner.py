"""Process-per-node blocking runner (legacy building block; reference
engine/node_runner.py:33-174).

A ``NodeRunner`` hosts ``step_fn(state) -> state`` and ``msg_handler(state, msg) -> state`` in
its own OS process and is driven by commands: ``step``, ``start_auto(interval)``, ``stop_auto``,
``state``, ``stop``; ``send_message`` posts into its inbox (drained before every step/command).
The child blocks on ``multiprocessing.connection.wait`` over both pipes with a timeout equal to
the time until the next auto-step, instead of a 10 ms polling loop.
"""
from __future__ import annotations

import threading
import time
from multiprocessing.connection import wait as conn_wait
from typing import Any, Callable, Optional

import cloudpickle

from .actor.backends.process import process_context


def _runner_main(cmd_conn, inbox_conn, blob: bytes) -> None:
    step_fn, msg_handler, init_state = cloudpickle.loads(blob)
    state = init_state if init_state is not None else {}
    auto, interval, last = False, 0.0, time.monotonic()
    cmd_conn.send(("ready", None))  # handshake: the interpreter and the callables are loaded

    def drain() -> None:
        nonlocal state
        while inbox_conn.poll(0):
            try:
                state = msg_handler(state, cloudpickle.loads(inbox_conn.recv_bytes()))
            except (EOFError, OSError):
                return

    while True:
        timeout = None
        if auto:
            timeout = max(0.0, interval - (time.monotonic() - last))
        ready = conn_wait([cmd_conn, inbox_conn], timeout)
        if auto and (time.monotonic() - last) >= interval:
            drain()
            state = step_fn(state)
            last = time.monotonic()
        if inbox_conn in ready:
            drain()
        if cmd_conn not in ready:
            continue
        try:
            op, payload = cmd_conn.recv()
        except (EOFError, OSError):
            break
        drain()
        if op == "stop":
            cmd_conn.send(("stopped", None))
            break
        if op == "step":
            state = step_fn(state)
            cmd_conn.send(("step", None))
        elif op == "start_auto":
            auto, interval, last = True, float(payload), time.monotonic()
            cmd_conn.send(("start_auto", interval))
        elif op == "stop_auto":
            auto = False
            cmd_conn.send(("stop_auto", None))
        elif op == "state":
            cmd_conn.send(("state", cloudpickle.dumps(state)))
        else:
            cmd_conn.send(("error", f"unknown op {op}"))


class NodeRunner:
    """A blocking, process-per-node state machine (legacy building block, kept for parity with the reference).

    The child process holds a ``state`` dict and two user functions: ``step_fn(state) -> state`` and
    ``msg_handler(state, msg) -> state``.  The parent drives it with commands.

    Parameters
    ----------
    step_fn : callable
    msg_handler : callable
    init_state : dict, optional

    Notes
    -----
    ``start()`` spawns the child and waits for its handshake; ``step()`` runs one step (the inbox is drained first);
    ``start_auto(interval)`` / ``stop_auto()`` step periodically; ``send_message(msg)`` posts into the inbox;
    ``state()`` fetches a copy of the state; ``stop()`` ends the process.  New code should use
    :class:`~byzpy_b200.engine.node.decentralized.DecentralizedNode`.
    """

    def __init__(self, step_fn: Callable[[dict], dict], msg_handler: Callable[[dict, Any], dict], *,
                 init_state: Optional[dict] = None) -> None:
        ctx = process_context()      # fork-server with torch preloaded ("spawn" semantics, fast start)
        self._cmd, child_cmd = ctx.Pipe(duplex=True)
        child_inbox, self._inbox = ctx.Pipe(duplex=False)
        blob = cloudpickle.dumps((step_fn, msg_handler, init_state))
        self._proc = ctx.Process(target=_runner_main, args=(child_cmd, child_inbox, blob), daemon=True)
        self._children = (child_cmd, child_inbox)
        self._lock = threading.Lock()
        self._inbox_lock = threading.Lock()
        self._pump_thread: Optional[threading.Thread] = None
        self._pump_stop = threading.Event()

    def start(self, ready_timeout: float = 120.0) -> None:
        self._proc.start()
        for c in self._children:
            c.close()
        # a spawned interpreter may take seconds to import its dependencies; commands carry short
        # timeouts, so wait for the child's handshake once instead of padding every timeout
        if not self._cmd.poll(ready_timeout):
            raise TimeoutError("NodeRunner child did not come up")
        self._cmd.recv()

    def _command(self, op: str, payload: Any = None, timeout: float = 5.0):
        with self._lock:
            self._cmd.send((op, payload))
            if not self._cmd.poll(timeout):
                raise TimeoutError(f"NodeRunner command {op!r} timed out")
            return self._cmd.recv()

    def stop(self) -> None:
        try:
            self._command("stop")
        except Exception:
            pass
        self._proc.join(timeout=5)
        if self._proc.is_alive():
            self._proc.terminate()
        self._stop_pump_thread()

    def step(self) -> None:
        self._command("step")

    def start_auto(self, interval_sec: float) -> None:
        self._command("start_auto", interval_sec)

    def stop_auto(self) -> None:
        self._command("stop_auto")

    def send_message(self, msg: Any) -> None:
        with self._inbox_lock:
            self._inbox.send_bytes(cloudpickle.dumps(msg))

    def state(self) -> dict:
        _, blob = self._command("state")
        return cloudpickle.loads(blob)

    def pump_once(self) -> None:
        try:
            self._command("step", timeout=1.0)
        except TimeoutError:
            pass

    def start_async(self, interval_sec: float = 0.01) -> None:
        if self._pump_thread is not None and self._pump_thread.is_alive():
            return
        self._pump_stop.clear()

        def loop() -> None:
            while not self._pump_stop.is_set():
                self.pump_once()
                self._pump_stop.wait(interval_sec)

        self._pump_thread = threading.Thread(target=loop, daemon=True)
        self._pump_thread.start()

    def _stop_pump_thread(self) -> None:
        self._pump_stop.set()
        if self._pump_thread is not None and self._pump_thread.is_alive():
            self._pump_thread.join(timeout=1.0)
        self._pump_thread = None


__all__ = ["NodeRunner"]
