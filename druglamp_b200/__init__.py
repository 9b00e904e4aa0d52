"""druglamp_b200 -- B200-native (sm_100a) implementation of DrugLAMP's cross-modal fusion +
contrastive hot path behind the reference's own nn.Module API.

The arithmetic lives in ``csrc/`` (hand-written CUDA behind the C ABI declared in
``include/druglamp_sm100.h``); this package is the host-side mirror of the reference interface
(``modules.py``, ``models.py``, ``ssl.py``).  There is no CPU fallback: using an op without the
built library, or on CPU tensors, raises.
"""
__version__ = "0.1.0"

from .kernels import compute_dtype, set_compute_dtype  # noqa: F401
from .modules import set_gcn_precision  # noqa: F401


def patch_reference() -> None:
    """Rebind the reference's hot-path classes to the sm_100a implementations so that the
    reference's own ``model/DrugLAMP*.py`` and ``trainer.py`` run unchanged on top of them.
    Call after ``/root/reference`` (or a checkout of Lzcstan/DrugLAMP) is importable and before
    the model is constructed.  See INTEGRATION.md."""
    import importlib

    from . import modules as M
    from . import ssl as S
    M.RETURN_CALLER_DTYPE = True      # the reference's own fp32 layers sit between the replaced modules
    bm = importlib.import_module("model.basic_model")
    bm.MolecularGCN = M.MolecularGCN
    bm.GuidedCrossAttention = M.GuidedCrossAttention
    bm.MultiHeadLinearAttention = M.MultiHeadLinearAttention
    bm.PairedMultimodelAttention = M.PairedMultimodelAttention
    bm.CrossModality = M.CrossModality
    bm.SSL = S.SSL
    bm.binary_cross_entropy = M.binary_cross_entropy
