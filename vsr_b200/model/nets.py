from ..nets import BaseNet, DRFNet, DRFSISRNet, SRFBNet  # noqa: F401
from ..edsr import EDSRNet  # noqa: F401
from ..duf import DUFNet  # noqa: F401
from ..rbpn import RBPNet  # noqa: F401
from ..frvsr import FRVSRNet  # noqa: F401
from ..toflow import TOFlowNet  # noqa: F401
