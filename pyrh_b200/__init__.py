"""pyrh_b200 -- B200-native (sm_100a CUDA, FP64) hot path of RH / pyrh.

Host-side mirror of the reference's operator interface for the 1-D formal
solution + LTE line opacity path; all compute goes through the C-ABI shared
library ``pyrh_b200/csrc/librhb200.so`` (include/rhb200.h).  There is no CPU
fallback: importing compute entry points without the CUDA library raises.
"""
from .linelist import LineTable  # noqa: F401

__version__ = "0.1.0"
