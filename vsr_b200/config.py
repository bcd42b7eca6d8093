"""Config-driven construction with the reference's semantics (src/main.py:56-76,167-178)."""
import torch

from . import model


def get_instance(module, config, *args):
    """getattr(module, config['name'])(*args, **config.get('kwargs', {}))  (main.py:167-178)."""
    cls = getattr(module, config["name"])
    kwargs = config.get("kwargs")
    return cls(*args, **kwargs) if kwargs else cls(*args)


def build_net(config):
    return get_instance(model.nets, config["net"])


def build_losses(config):
    """loss in torch.nn first, else in the model.losses namespace (main.py:59-67)."""
    fns, weights = [], []
    for c in config["losses"]:
        mod = torch.nn if c["name"] in dir(torch.nn) else model.losses
        fns.append(get_instance(mod, c))
        weights.append(c["weight"])
    return fns, weights


def build_metrics(config):
    return [get_instance(model.metrics, c) for c in config["metrics"]]


def build_optimizer(config, net):
    return get_instance(torch.optim, config["optimizer"], net.parameters())
