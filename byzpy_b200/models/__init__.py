"""Model zoo for the target configs (BASELINE.json): SmallCNN (MNIST),
ResNet-18/34/50, BERT-base.  All are random-init ``torch.nn`` modules whose
parameters get re-homed into flat arenas by :mod:`byzpy_b200.parallel.arena`."""
from .bert import BertConfig, BertEncoder, BertForMaskedLM, bert_base
from .resnet import ResNet, resnet18, resnet34, resnet50
from .smallcnn import SmallCNN

_REGISTRY = {
    "smallcnn": SmallCNN,
    "resnet18": resnet18,
    "resnet34": resnet34,
    "resnet50": resnet50,
    "bert-base": bert_base,
    "bert_base": bert_base,
}


def build_model(name: str, **kw):
    """Model by name: ``smallcnn``, ``resnet18`` / ``34`` / ``50``, ``bert-base``; keyword arguments go to the
    constructor.

    >>> from byzpy_b200.models import build_model
    >>> type(build_model("smallcnn")).__name__
    'SmallCNN'
    """
    try:
        return _REGISTRY[name.lower()](**kw)
    except KeyError as exc:
        raise ValueError(f"unknown model {name!r}; choose from {sorted(_REGISTRY)}") from exc


__all__ = ["SmallCNN", "ResNet", "resnet18", "resnet34", "resnet50", "BertConfig", "BertEncoder",
           "BertForMaskedLM", "bert_base", "build_model"]
