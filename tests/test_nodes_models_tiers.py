"""Node classes on the host path, model zoo and checkpoint tiers (CPU): device nodes used as
ordinary nodes, distributed nodes' input routing, BERT / SmallCNN forward+backward, model registry,
checkpoint resume details (mirrors reference tests/engine/node/test_distributed.py and the example
node classes of reference examples/ps/nodes.py)."""
import asyncio

import pytest
import torch
import torch.nn as nn

from byzpy_b200.aggregators.coordinate_wise import CoordinateWiseMedian, CoordinateWiseTrimmedMean
from byzpy_b200.attacks import EmpireAttack, LabelFlipAttack, LittleAttack, SignFlipAttack
from byzpy_b200.engine.graph.pool import ActorPool, ActorPoolConfig
from byzpy_b200.engine.node.device import (DeviceByzantineNode, DeviceHonestNode, DeviceP2PByzantineNode,
                                           DeviceP2PHonestNode)
from byzpy_b200.engine.node.distributed import DistributedByzantineNode, DistributedHonestNode
from byzpy_b200.engine.parameter_server.ps import ParameterServer
from byzpy_b200.models import BertConfig, BertEncoder, BertForMaskedLM, SmallCNN, bert_base, build_model
from byzpy_b200.parallel.arena import flatten_grads, flatten_params
from byzpy_b200.utils.checkpoint import FORMAT, load_checkpoint, load_reference_state_dict, save_checkpoint


def run(coro):
    return asyncio.run(coro)


def _src(seed=0, n=8, d=6, k=3):
    g = torch.Generator().manual_seed(seed)

    def nxt():
        return torch.randn(n, d, generator=g), torch.randint(0, k, (n,), generator=g)

    return nxt


# ------------------------------------------------------------------------------- device nodes on CPU
def test_device_honest_node_is_an_ordinary_node():
    torch.manual_seed(0)
    node = DeviceHonestNode(nn.Linear(6, 3), data=_src(), lr=0.1, momentum=0.0, device="cpu", name="h0")
    assert node.device.type == "cpu" and node.worker.role == "honest" and node.worker.name == "h0"
    x, y = node.next_batch()
    g = node.honest_gradient(x, y)
    loss = nn.functional.cross_entropy(node.model(x), y)
    exp = torch.cat([t.reshape(-1) for t in torch.autograd.grad(loss, list(node.model.parameters()))])
    assert torch.allclose(g, exp, atol=1e-6) and g.numel() == 21
    before = flatten_params(node.model).clone()
    node.apply_server_gradient(g)
    assert torch.allclose(flatten_params(node.model), before - 0.1 * g, atol=1e-6)
    sd = node.dump_state_dict()
    assert set(sd) == {"weight", "bias"} and all(v.device.type == "cpu" and not v.requires_grad for v in sd.values())
    nn.Linear(6, 3).load_state_dict(sd, strict=True)
    with pytest.raises(RuntimeError, match="no data source"):
        DeviceHonestNode(nn.Linear(6, 3), device="cpu").next_batch()


def test_device_honest_node_preprocess_is_applied():
    node = DeviceHonestNode(nn.Linear(6, 3), data=_src(), device="cpu", preprocess=lambda x: x * 0.0)
    x, y = node.next_batch()
    g = node.honest_gradient(x, y)
    assert torch.count_nonzero(g[:18]) == 0 and torch.count_nonzero(g[18:]) > 0     # zero inputs: only bias grads


def test_device_byzantine_node_routes_attack_inputs():
    vs = [torch.ones(21), 3 * torch.ones(21)]
    omni = DeviceByzantineNode(EmpireAttack(scale=-1.0), device="cpu")
    assert omni.worker is None and omni.model is None and omni.fold(2).kind == "virtual"
    x, y = omni.next_batch()
    assert x.numel() == 0 and y.dtype == torch.long
    assert torch.allclose(omni.byzantine_gradient(x, y, honest_grads=vs), torch.full((21,), -2.0))
    omni.apply_server_gradient(torch.zeros(21))                      # no model: silently ignored
    with pytest.raises(ValueError, match="needs the node's own model"):
        DeviceByzantineNode(SignFlipAttack(), device="cpu")
    torch.manual_seed(1)
    flip = DeviceByzantineNode(SignFlipAttack(scale=-2.0), model=nn.Linear(6, 3), data=_src(3), device="cpu")
    assert flip.worker is not None and flip.worker.role == "byzantine" and flip.fold(0).kind == "scale"
    out = flip.byzantine_gradient(torch.empty(0), torch.empty(0, dtype=torch.long), honest_grads=vs)
    assert torch.allclose(out, -2.0 * flatten_grads(flip.model))
    before = flatten_params(flip.model).clone()
    flip.apply_server_gradient(torch.ones(21))
    assert not torch.equal(flatten_params(flip.model), before)
    lab = DeviceByzantineNode(LabelFlipAttack(num_classes=3), model=nn.Linear(6, 3), data=_src(4), device="cpu")
    assert lab.byzantine_gradient(torch.empty(0), torch.empty(0, dtype=torch.long)).numel() == 21


def test_device_nodes_train_through_the_generic_round():
    torch.manual_seed(0)
    w = torch.randn(6, 3)

    def src(seed):
        g = torch.Generator().manual_seed(seed)

        def nxt():
            x = torch.randn(64, 6, generator=g)
            return x, (x @ w).argmax(1)

        return nxt

    hon = [DeviceHonestNode(nn.Linear(6, 3), data=src(i), lr=0.5, momentum=0.0, device="cpu") for i in range(5)]
    for h in hon[1:]:
        h.model.load_state_dict(hon[0].model.state_dict())
    byz = [DeviceByzantineNode(LittleAttack(f=1), device="cpu")]
    ps = ParameterServer(hon, byz, CoordinateWiseTrimmedMean(f=1))
    x, y = src(99)()
    first = nn.functional.cross_entropy(hon[0].model(x), y).item()
    for _ in range(40):
        ps.round_sync()
    last = nn.functional.cross_entropy(hon[0].model(x), y).item()
    assert last < 0.6 * first and ps.rounds == 40
    assert all(torch.equal(flatten_params(h.model), flatten_params(hon[0].model)) for h in hon)


def test_p2p_device_nodes_expose_the_mixin_contract():
    node = DeviceP2PHonestNode(nn.Linear(6, 3), CoordinateWiseMedian(), data=_src(), device="cpu",
                               preprocess=lambda x: x + 1.0, name="p")
    x, y = node.next_batch()
    assert x.shape == (8, 6) and node.p2p_pre is None and node.name == "p" and set(node.dump_state_dict()) == {"weight", "bias"}
    half = node.p2p_half_step(lr=0.1) if hasattr(node, "p2p_half_step") else None
    assert half is None or half.numel() == 21
    b = DeviceP2PByzantineNode(EmpireAttack(), device="cpu")
    assert b.device.type == "cpu" and b.name == "p2p-byz" and isinstance(b.attack, EmpireAttack)


# --------------------------------------------------------------------------------- distributed nodes
class _Hon(DistributedHonestNode):
    def __init__(self, **kw):
        super().__init__(actor_pool=[ActorPoolConfig("thread")], aggregator=CoordinateWiseMedian(), **kw)
        self.applied = []

    def next_batch(self):
        return torch.ones(2), torch.zeros(2)

    def local_honest_gradient(self, *, x, y):
        return x * 3 + y

    def apply_server_gradient(self, g):
        self.applied.append(g)


class _Byz(DistributedByzantineNode):
    def next_batch(self):
        return torch.empty(0), torch.empty(0, dtype=torch.long)

    def apply_server_gradient(self, g):
        pass


def test_distributed_honest_node_wiring():
    h = _Hon(name="alpha", metadata={"k": 1})
    assert h.application.name == "alpha" and isinstance(h.pool, ActorPool) and h.pool is h.application.pool
    assert set(h.application.list_pipelines()) == {"aggregate", "honest_gradient"}
    assert torch.equal(h.honest_gradient(torch.ones(2), torch.ones(2)), torch.full((2,), 4.0))
    assert _Hon().application.name == "_Hon"                         # default name = class name
    with pytest.raises(NotImplementedError):
        DistributedHonestNode.local_honest_gradient(h, x=1, y=2)
    run(h.shutdown_distributed())


@pytest.mark.parametrize("attack,kwargs,missing", [
    (EmpireAttack(), dict(honest_grads=[torch.ones(2)]), "honest_grads"),
    (SignFlipAttack(), dict(base_grad=torch.ones(2)), "base_grad"),
    (LabelFlipAttack(num_classes=2), dict(model=nn.Linear(2, 2), x=torch.ones(1, 2), y=torch.zeros(1, dtype=torch.long)), "model"),
])
def test_distributed_byzantine_input_routing(attack, kwargs, missing):
    b = _Byz(actor_pool=[ActorPoolConfig("thread")], attack=attack)
    routed = b.prepare_attack_inputs(**kwargs)
    assert set(routed) == set(kwargs) and list(routed) == b._attack_keys()
    with pytest.raises(ValueError, match=missing):
        b.prepare_attack_inputs(**{k: v for k, v in kwargs.items() if k != missing})
    out = run(b.byzantine_gradient_async(**kwargs))
    assert isinstance(out, torch.Tensor)
    assert torch.equal(run(b.run_attack(inputs=routed)), out) or isinstance(attack, LabelFlipAttack)
    run(b.shutdown_distributed())


def test_distributed_byzantine_label_flip_needs_a_batch():
    b = _Byz(actor_pool=[ActorPoolConfig("thread")], attack=LabelFlipAttack(num_classes=2))
    with pytest.raises(ValueError, match="'x' and 'y'"):
        b.prepare_attack_inputs(model=nn.Linear(2, 2))
    run(b.shutdown_distributed())


def test_distributed_byzantine_custom_override_signature():
    class Custom(_Byz):
        def __init__(self):
            super().__init__(actor_pool=[ActorPoolConfig("thread")])

        def byzantine_gradient(self, x, y, honest_grads=None, boost=2.0):
            return -boost * torch.stack(list(honest_grads)).mean(0)

    c = Custom()
    assert c._custom_input_keys == ("x", "y", "honest_grads", "boost") and c._custom_required_keys == ("x", "y")
    g = [torch.ones(2), 3 * torch.ones(2)]
    assert torch.equal(c.byzantine_gradient_for_next_batch(g), torch.full((2,), -4.0))
    with pytest.raises(ValueError, match="'x'"):
        c._build_custom_inputs(y=1)
    assert set(c._build_custom_inputs(x=1, y=2, honest_grads=None)) == {"x", "y"}

    class Inherits(Custom):        # the override is inherited, not lost
        pass

    assert Inherits()._custom_bz_callable is not None
    run(c.shutdown_distributed())


# ---------------------------------------------------------------------------------------- model zoo
def test_small_cnn_shapes_names_and_gradients():
    m = SmallCNN()
    assert [n for n, _ in m.named_parameters()] == ["conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias",
                                                    "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"]
    assert sum(p.numel() for p in m.parameters()) == 421_642
    out = m(torch.randn(4, 1, 28, 28))
    assert out.shape == (4, 10)
    out.sum().backward()
    assert all(p.grad is not None for p in m.parameters())
    assert SmallCNN(in_channels=3, num_classes=7)(torch.randn(2, 3, 28, 28)).shape == (2, 7)


def test_bert_tiny_forward_backward_masking_and_tied_head():
    torch.manual_seed(0)
    cfg = dict(vocab_size=97, hidden=32, layers=2, heads=4, ffn=64, max_pos=16)
    m = bert_base(**cfg)
    assert isinstance(m, BertForMaskedLM) and isinstance(m.bert, BertEncoder) and m.bert.config == BertConfig(**cfg)
    ids = torch.randint(0, 97, (3, 10))
    logits = m(ids)
    assert logits.shape == (3, 10, 97)
    nn.functional.cross_entropy(logits.reshape(-1, 97), ids.reshape(-1)).backward()
    assert m.bert.tok.weight.grad is not None and m.bias.grad is not None
    assert "decoder.weight" not in dict(m.named_parameters())        # the head reuses the embedding matrix
    # padding positions must not influence the unpadded positions
    m.eval()
    mask = torch.ones(1, 10, dtype=torch.long)
    mask[0, 6:] = 0
    a = m(ids[:1], attention_mask=mask)
    ids2 = ids[:1].clone()
    ids2[0, 6:] = 5
    b = m(ids2, attention_mask=mask)
    assert torch.allclose(a[0, :6], b[0, :6], atol=1e-5) and not torch.allclose(a[0, 6:], b[0, 6:], atol=1e-5)
    h = m.bert(ids[:1], token_type=torch.ones(1, 10, dtype=torch.long))
    assert h.shape == (1, 10, 32) and not torch.allclose(h, m.bert(ids[:1]))


def test_bert_base_parameter_count_without_allocating():
    with torch.device("meta"):
        m = BertForMaskedLM(BertConfig())
    n = sum(p.numel() for p in m.parameters())
    assert 109_000_000 < n < 111_000_000 and len(m.bert.blocks) == 12


def test_model_registry():
    assert isinstance(build_model("smallcnn"), SmallCNN) and isinstance(build_model("SmallCNN", num_classes=3), SmallCNN)
    assert build_model("bert-base", vocab_size=11, hidden=8, layers=1, heads=2, ffn=8, max_pos=4).bert.config.layers == 1
    with pytest.raises(ValueError, match="unknown model"):
        build_model("vgg")


# --------------------------------------------------------------------------------------- checkpoints
def _build_ps(seed=0):
    torch.manual_seed(seed)
    hon = [DeviceHonestNode(nn.Linear(6, 3), data=_src(i), lr=0.1, momentum=0.9, device="cpu", name=f"h{i}")
           for i in range(3)]
    byz = [DeviceByzantineNode(SignFlipAttack(), model=nn.Linear(6, 3), data=_src(9), device="cpu", name="b0"),
           DeviceByzantineNode(EmpireAttack(), device="cpu", name="b1")]
    return ParameterServer(hon, byz, CoordinateWiseMedian(), update_byzantines=True)


def test_checkpoint_contents_and_resume_continues_identically(tmp_path):
    ps = _build_ps()
    for _ in range(3):
        ps.round_sync()
    path = str(tmp_path / "ck.pt")
    save_checkpoint(path, ps, extra={"note": "hi"})
    blob = torch.load(path, weights_only=False)
    assert blob["format"] == FORMAT and blob["round"] == 3 and blob["extra"] == {"note": "hi"} and "torch" in blob["rng"]
    assert [(r["name"], r["role"]) for r in blob["nodes"]] == [("h0", "honest"), ("h1", "honest"), ("h2", "honest"),
                                                               ("b0", "byzantine")]      # model-less nodes carry no state
    assert blob["nodes"][0]["optimizer"] is not None and blob["nodes"][0]["momentum"] is None
    ps2 = _build_ps(seed=5)                                          # different init, same topology
    assert load_checkpoint(path, ps2) == 3 and ps2.rounds == 3
    for a, b in zip(ps.hon + ps.byz[:1], ps2.hon + ps2.byz[:1]):
        assert torch.equal(flatten_params(a.model), flatten_params(b.model))
        a.data, b.data = _src(77), _src(77)                          # same future batches on both sides
    ps.round_sync()
    ps2.round_sync()
    for a, b in zip(ps.hon, ps2.hon):                                # momentum buffers were restored too
        assert torch.allclose(flatten_params(a.model), flatten_params(b.model), atol=1e-7)


def test_checkpoint_round_override_rng_and_errors(tmp_path):
    ps = _build_ps()
    path = str(tmp_path / "ck.pt")
    save_checkpoint(path, ps, round_index=42)
    torch.manual_seed(123)
    expected = torch.rand(3)
    torch.manual_seed(123)
    save_checkpoint(path, ps, round_index=42)
    torch.rand(10)                                                   # advance the generator
    assert load_checkpoint(path, ps) == 42
    assert torch.equal(torch.rand(3), expected)                      # RNG state was restored
    torch.rand(10)
    state = torch.get_rng_state()
    load_checkpoint(path, ps, restore_rng=False)
    assert torch.equal(torch.get_rng_state(), state)
    small = ParameterServer(ps.hon[:2], [], CoordinateWiseMedian())
    with pytest.raises(ValueError, match="different number of nodes"):
        load_checkpoint(path, small)
    bad = str(tmp_path / "bad.pt")
    torch.save({"format": "something-else"}, bad)
    with pytest.raises(ValueError, match="not a"):
        load_checkpoint(bad, ps)


def test_reference_state_dict_loader_is_strict_by_default():
    src, dst = nn.Linear(4, 2), nn.Linear(4, 2)
    snap = {k: v.detach().cpu() for k, v in src.state_dict().items()}
    load_reference_state_dict(dst, snap)
    assert torch.equal(dst.weight, src.weight)
    with pytest.raises(RuntimeError):
        load_reference_state_dict(dst, {"weight": snap["weight"]})
    res = load_reference_state_dict(dst, {"weight": snap["weight"]}, strict=False)
    assert res.missing_keys == ["bias"]
