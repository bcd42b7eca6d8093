"""Stub loader for the real reference modules.

`import src` of the reference fails here (nibabel / box / SimpleITK / imageio are absent and
src/model/nets/__init__.py imports the un-importable EDVR net), so the package __init__ files are
replaced by empty modules and the needed files are executed in place from /root/reference.
Reference sources are never copied into this repository.
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("VSR_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "model", "nets", "drf_net.py"))


def _stub(name, path):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    return sys.modules[name]


def _load(name, relpath):
    if name in sys.modules and getattr(sys.modules[name], "__file__", None):
        return sys.modules[name]
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """Returns a namespace with the reference classes on the hot path."""
    if not available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT}")
    src = os.path.join(REFERENCE_ROOT, "src")
    _stub("src", src)
    _stub("src.model", os.path.join(src, "model"))
    _stub("src.model.nets", os.path.join(src, "model", "nets"))
    _load("src.model.nets.base_net", "src/model/nets/base_net.py")
    ns = types.SimpleNamespace()
    ns.DRFNet = _load("src.model.nets.drf_net", "src/model/nets/drf_net.py").DRFNet
    ns.DRFSISRNet = _load("src.model.nets.drf_sisr_net", "src/model/nets/drf_sisr_net.py").DRFSISRNet
    ns.SRFBNet = _load("src.model.nets.srfb_net", "src/model/nets/srfb_net.py").SRFBNet
    ns.EDSRNet = _load("src.model.nets.edsr_net", "src/model/nets/edsr_net.py").EDSRNet
    ns.RBPNet = _load("src.model.nets.rbp_net", "src/model/nets/rbp_net.py").RBPNet
    ns.FRVSRNet = _load("src.model.nets.frvsr_net", "src/model/nets/frvsr_net.py").FRVSRNet
    ns.TOFlowNet = _load("src.model.nets.toflow_net", "src/model/nets/toflow_net.py").TOFlowNet
    losses = _load("src.model.losses", "src/model/losses.py")
    metrics = _load("src.model.metrics", "src/model/metrics.py")
    utils = _load("src.utils", "src/utils.py")
    ns.HuberLoss, ns.CharbonnierLoss = losses.HuberLoss, losses.CharbonnierLoss
    ns.PSNR, ns.SSIM = metrics.PSNR, metrics.SSIM
    ns.CardiacPSNR, ns.CardiacSSIM = metrics.CardiacPSNR, metrics.CardiacSSIM
    ns.denormalize = utils.denormalize
    return ns
