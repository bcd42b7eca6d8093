from ..losses import CharbonnierLoss, FusedL1Loss, FusedMSELoss, HuberLoss  # noqa: F401
