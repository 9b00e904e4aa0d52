"""druglamp_b200 -- B200-native (sm_100a) implementation of DrugLAMP's cross-modal
fusion + contrastive hot path behind the reference's own nn.Module API.

The arithmetic lives in ``csrc/`` (hand-written CUDA behind the C ABI declared in
``include/druglamp_sm100.h``); this package is the host-side mirror of the reference
interface.  There is no CPU fallback: using an op without the built library raises.
"""
__version__ = "0.1.0"
