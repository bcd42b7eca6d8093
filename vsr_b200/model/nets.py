from ..nets import BaseNet, DRFNet, DRFSISRNet  # noqa: F401
