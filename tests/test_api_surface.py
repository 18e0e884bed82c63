"""Public API surface against the reference, module by module: every module path of the reference package imports
through the ``byzpy`` alias, and every public name it defines (``__all__``, else its top-level classes / functions)
exists here.  Skipped when the reference tree is not on the box."""
import ast
import importlib
import os

import pytest

REF = "/root/reference/python/byzpy"


def _reference_modules():
    for dp, dn, fn in os.walk(REF):
        dn[:] = [d for d in dn if d not in ("tests", "__pycache__")]
        for f in fn:
            if not f.endswith(".py"):
                continue
            path = os.path.join(dp, f)
            rel = os.path.relpath(path, REF)[:-3].replace(os.sep, ".")
            if rel.endswith("__init__"):
                rel = rel[:-9].rstrip(".")
            yield "byzpy" + ("." + rel if rel else ""), path


def _public_names(path):
    tree = ast.parse(open(path, encoding="utf-8").read())
    names, declared = set(), None
    for n in tree.body:
        if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)) and not n.name.startswith("_"):
            names.add(n.name)
        if isinstance(n, ast.Assign) and any(isinstance(t, ast.Name) and t.id == "__all__" for t in n.targets):
            try:
                declared = set(ast.literal_eval(n.value))
            except Exception:
                pass
    return declared if declared is not None else names


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not available")
def test_every_reference_module_and_public_name_exists_here(byzpy_alias):
    modules = list(_reference_modules())
    assert len(modules) > 80
    missing, checked = [], 0
    for name, path in modules:
        mod = importlib.import_module(name)              # the module path itself
        for sym in sorted(_public_names(path)):
            checked += 1
            if not hasattr(mod, sym):
                missing.append(f"{name}.{sym}")
    assert checked > 150 and not missing, missing


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not available")
def test_public_methods_and_constructor_parameters_of_every_reference_class_exist_here(byzpy_alias):
    import inspect

    missing, classes = [], 0
    for name, path in _reference_modules():
        mod = importlib.import_module(name)
        for node in ast.parse(open(path, encoding="utf-8").read()).body:
            if not isinstance(node, ast.ClassDef) or node.name.startswith("_") or not hasattr(mod, node.name):
                continue
            ours = getattr(mod, node.name)
            classes += 1
            for m in node.body:
                if not isinstance(m, (ast.FunctionDef, ast.AsyncFunctionDef)):
                    continue
                if m.name == "__init__":
                    want = [a.arg for a in m.args.args[1:]] + [a.arg for a in m.args.kwonlyargs]
                    sig = inspect.signature(ours.__init__)
                    if any(p.kind is p.VAR_KEYWORD for p in sig.parameters.values()):
                        continue
                    missing += [f"{name}.{node.name}(…{p}=)" for p in want if p not in sig.parameters]
                elif not m.name.startswith("_") and not hasattr(ours, m.name):
                    missing.append(f"{name}.{node.name}.{m.name}")
    assert classes > 90 and not missing, missing


def test_ucx_transport_listener_endpoint_control_and_payload_helpers():
    """The reference's UCX helper surface (create_listener / create_endpoint / send_control / recv_control /
    send_payload / recv_payload, reference transports/ucx.py:84-277) on stream endpoints."""
    import asyncio

    import torch

    from byzpy_b200.engine.actor.channels import Endpoint
    from byzpy_b200.engine.actor.transports import ucx

    async def scenario():
        got = asyncio.get_running_loop().create_future()

        async def serve(ep):
            ctrl = await ucx.recv_control(ep)
            payload = await ucx.recv_payload(ep)
            await ucx.send_control(ep, {"ack": ctrl["seq"]})
            got.set_result((ctrl, payload))

        server = await ucx.create_listener(serve, host="127.0.0.1", port=0)
        port = server.sockets[0].getsockname()[1]
        ep = await ucx.create_endpoint("127.0.0.1", port)
        assert ep is not await ucx.get_endpoint("127.0.0.1", port)            # dedicated, not the pooled one
        obj = {"v": torch.arange(5.0), "ep": Endpoint("ucx", f"127.0.0.1:{port}", "a")}
        await ucx.send_control(ep, {"seq": 7})
        tag, desc = ucx.pack_payload(obj)
        await ucx.send_payload(ep, tag, desc, obj)
        assert (await ucx.recv_control(ep)) == {"ack": 7}
        ctrl, payload = await asyncio.wait_for(got, 5.0)
        assert ctrl == {"seq": 7} and torch.equal(payload["v"], obj["v"]) and payload["ep"] == obj["ep"]
        ep[1].close()
        await ucx.clear_pool()
        server.close()
        await asyncio.wait_for(server.wait_closed(), 2.0)

    asyncio.run(scenario())


def test_dependency_hooks():
    from byzpy_b200 import _dependencies as d

    assert d.get_gpu_optional_dependencies() == [] and any(x.startswith("torch") for x in d.get_dependencies())
    assert any("pytest" in x for x in d.get_dev_optional_dependencies())


def _ref_params(fn_node, drop_first):
    a = fn_node.args
    pos = [x.arg for x in a.posonlyargs + a.args]
    if drop_first and pos and pos[0] in ("self", "cls"):
        pos = pos[1:]
    return pos, [x.arg for x in a.kwonlyargs]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not available")
def test_parameters_of_every_public_function_and_method_of_the_reference_are_accepted_here(byzpy_alias):
    """Every parameter NAME of the reference's public functions and methods exists here (extra parameters with
    defaults are fine), and the shared positional parameters come in the same order."""
    import inspect

    diffs, compared = [], 0

    def compare(qual, node, ours, drop_first):
        nonlocal compared
        try:
            sig = inspect.signature(ours)
        except (TypeError, ValueError):
            return
        compared += 1
        pos, kwo = _ref_params(node, drop_first)
        ps = list(sig.parameters.values())
        if drop_first and ps and ps[0].name in ("self", "cls"):
            ps = ps[1:]
        names = [p.name for p in ps]
        var_kw = any(p.kind is p.VAR_KEYWORD for p in ps)
        var_pos = any(p.kind is p.VAR_POSITIONAL for p in ps)
        lacking = [p for p in pos + kwo if p not in names]
        if lacking and not var_kw and not (var_pos and not kwo):
            diffs.append((qual, "missing", lacking))
            return
        ours_pos = [p.name for p in ps if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
        common = [p for p in pos if p in ours_pos]
        if not var_pos and [p for p in ours_pos if p in common] != common:
            diffs.append((qual, "order", pos, ours_pos))

    for name, path in _reference_modules():
        mod = importlib.import_module(name)
        for node in ast.parse(open(path, encoding="utf-8").read()).body:
            if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef)) and not node.name.startswith("_") \
                    and hasattr(mod, node.name):
                compare(f"{name}.{node.name}", node, getattr(mod, node.name), False)
            if isinstance(node, ast.ClassDef) and not node.name.startswith("_") and hasattr(mod, node.name):
                cls = getattr(mod, node.name)
                for m in node.body:
                    if not isinstance(m, (ast.FunctionDef, ast.AsyncFunctionDef)) or m.name.startswith("_") \
                            or not hasattr(cls, m.name):
                        continue
                    raw = inspect.getattr_static(cls, m.name)
                    if isinstance(raw, property) or any(isinstance(d, ast.Name) and d.id == "property"
                                                        for d in m.decorator_list):
                        continue
                    bound = inspect.ismethod(getattr(cls, m.name))           # classmethod: cls already bound
                    is_static = isinstance(raw, staticmethod)
                    ref_is_plain = not any(isinstance(d, ast.Name) and d.id in ("staticmethod",)
                                           for d in m.decorator_list)
                    node2 = m
                    if ref_is_plain and (bound or is_static):
                        node2 = ast.FunctionDef(name=m.name, args=ast.arguments(
                            posonlyargs=[], args=m.args.args[1:], vararg=m.args.vararg, kwonlyargs=m.args.kwonlyargs,
                            kw_defaults=m.args.kw_defaults, kwarg=m.args.kwarg, defaults=[]), body=[], decorator_list=[])
                    compare(f"{name}.{node.name}.{m.name}", node2, getattr(cls, m.name),
                            not (bound or is_static) and ref_is_plain)
    assert compared > 300 and not diffs, diffs


def test_low_level_helpers_accept_the_reference_calling_conventions():
    import asyncio

    import torch

    from byzpy_b200.configs.backend import get_backend, set_backend, use_backend
    from byzpy_b200.engine.actor.backends.remote import RemoteActorServer
    from byzpy_b200.engine.actor.channels import Endpoint
    from byzpy_b200.engine.actor.transports import tcp

    set_backend(backend="numpy")
    assert get_backend().name == "numpy"
    with use_backend(backend="pytorch"):
        assert get_backend().name == "torch"
    set_backend(get_backend())                       # a backend object is accepted too
    set_backend("torch")
    with pytest.raises(ValueError):
        set_backend("jax")

    async def scenario():
        srv = RemoteActorServer("127.0.0.1", 0)
        await srv.start()
        try:
            ep = Endpoint("tcp", f"127.0.0.1:{srv.port}", "box-owner")
            # the reference's convention: address string + keyword endpoints
            await tcp.chan_put(f"127.0.0.1:{srv.port}", from_ep=None, to_ep=ep, name="inbox", payload={"x": torch.ones(2)})
            got = await tcp.chan_get(f"127.0.0.1:{srv.port}", name="inbox", timeout=2.0, actor_id="box-owner")
            assert torch.equal(got["x"], torch.ones(2))
            # this package's convention
            await tcp.chan_put("127.0.0.1", srv.port, "box-owner", "inbox", 5)
            assert await tcp.chan_get("127.0.0.1", srv.port, "box-owner", "inbox", 2.0) == 5
        finally:
            await srv.stop()

    asyncio.run(scenario())


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not available")
def test_class_level_attributes_of_the_reference_exist_with_the_same_literal_values(byzpy_alias):
    """``name``, ``supports_subtasks``, ``input_key``, ``max_subtasks_inflight`` ...: every public class attribute the
    reference assigns in a class body exists here; where it is a literal, the value is the same."""
    problems, seen = [], 0
    dataclass_fields = {("SubTask", "args"), ("SubTask", "kwargs"), ("GraphNode", "inputs")}      # default_factory fields
    for name, path in _reference_modules():
        mod = importlib.import_module(name)
        for node in ast.parse(open(path, encoding="utf-8").read()).body:
            if not isinstance(node, ast.ClassDef) or node.name.startswith("_") or not hasattr(mod, node.name):
                continue
            cls = getattr(mod, node.name)
            for m in node.body:
                if isinstance(m, ast.Assign):
                    targets, value = [t.id for t in m.targets if isinstance(t, ast.Name)], m.value
                elif isinstance(m, ast.AnnAssign) and isinstance(m.target, ast.Name) and m.value is not None:
                    targets, value = [m.target.id], m.value
                else:
                    continue
                for t in targets:
                    if t.startswith("_") or (node.name, t) in dataclass_fields:
                        continue
                    seen += 1
                    if not hasattr(cls, t):
                        problems.append(f"{name}.{node.name}.{t} missing")
                        continue
                    try:
                        want = ast.literal_eval(value)
                    except (ValueError, SyntaxError):
                        continue
                    have = getattr(cls, t)
                    if not callable(have) and have != want:
                        problems.append(f"{name}.{node.name}.{t}: {have!r} != {want!r}")
    assert seen > 80 and not problems, problems


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not available")
def test_default_argument_values_match_the_reference(byzpy_alias):
    """Literal defaults of constructor / function / method parameters (chunk sizes, tolerances, iteration counts,
    ``eval_interval`` ...).  One deliberate difference: servers listen on the loopback interface by default (the
    reference binds 0.0.0.0 and unpickles whatever connects)."""
    import inspect

    allowed = {("byzpy.engine.actor.backends.gpu.UCXRemoteActorServer.__init__", "host"),
               ("byzpy.engine.actor.backends.gpu.start_ucx_actor_server", "host")}
    diffs, compared = [], 0

    def ref_defaults(fn):
        a = fn.args
        pos = a.posonlyargs + a.args
        pairs = list(zip(pos[len(pos) - len(a.defaults):], a.defaults)) + \
            [(k, d) for k, d in zip(a.kwonlyargs, a.kw_defaults) if d is not None]
        out = {}
        for arg, d in pairs:
            try:
                out[arg.arg] = ast.literal_eval(d)
            except (ValueError, SyntaxError):
                pass
        return out

    def compare(qual, node, ours):
        nonlocal compared
        try:
            sig = inspect.signature(ours)
        except (TypeError, ValueError):
            return
        for k, v in ref_defaults(node).items():
            p = sig.parameters.get(k)
            if p is None or (qual, k) in allowed:
                continue
            compared += 1
            if p.default is inspect.Parameter.empty:
                diffs.append(f"{qual}({k}): required here, default {v!r} in the reference")
            elif p.default != v:
                diffs.append(f"{qual}({k}={p.default!r}) != {v!r}")

    for name, path in _reference_modules():
        mod = importlib.import_module(name)
        for node in ast.parse(open(path, encoding="utf-8").read()).body:
            if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef)) and not node.name.startswith("_") \
                    and hasattr(mod, node.name):
                compare(f"{name}.{node.name}", node, getattr(mod, node.name))
            if isinstance(node, ast.ClassDef) and not node.name.startswith("_") and hasattr(mod, node.name):
                cls = getattr(mod, node.name)
                for m in node.body:
                    if isinstance(m, (ast.FunctionDef, ast.AsyncFunctionDef)) and hasattr(cls, m.name) \
                            and (m.name == "__init__" or not m.name.startswith("_")) \
                            and not isinstance(inspect.getattr_static(cls, m.name), property):
                        compare(f"{name}.{node.name}.{m.name}", m, getattr(cls, m.name))
    assert compared > 180 and not diffs, diffs
