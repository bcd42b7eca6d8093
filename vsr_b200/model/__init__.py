"""Namespace mirror of the reference's `src.model` so that config-driven construction
(`getattr(src.model.nets, config.net.name)(**config.net.kwargs)`, main.py:56,167-178) resolves to
the B200 implementations: vsr_b200.model.nets / .losses / .metrics."""
from . import losses, metrics, nets  # noqa: F401
