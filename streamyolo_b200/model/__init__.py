"""Host-side mirror of the reference's ``exps/model`` operator API:
same class names, constructor signatures, forward modes and state_dict keys; all arithmetic
runs in libstreamyolo_sm100.so."""
from .network_blocks import BaseConv, Bottleneck, CSPLayer, DWConv, Focus, SPPBottleneck  # noqa: F401
from .darknet import CSPDarknet  # noqa: F401
from .dfp_pafpn import DFPPAFPN  # noqa: F401
from .tal_head import TALHead  # noqa: F401
from .pipe_head import PIPEHead  # noqa: F401
from .yolox import YOLOX  # noqa: F401
