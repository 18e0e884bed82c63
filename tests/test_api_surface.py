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
def test_every_reference_module_and_public_name_exists_here():
    from byzpy_b200.compat import install_alias

    install_alias()
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
def test_public_methods_and_constructor_parameters_of_every_reference_class_exist_here():
    import inspect

    from byzpy_b200.compat import install_alias

    install_alias()
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
