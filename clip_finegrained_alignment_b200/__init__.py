"""B200-native (sm_100a) contrastive-loss + AdamSPD hot path of tpeat/clip-finegrained-alignment.

Drop-in for the reference's `finetune/losses.py` (SPARCLoss, CustomCLIPLoss, CLIPCountLoss, CountLoss) and
`finetune/optimizers.py` (AdamSPD): same constructors, call signatures and result dicts, with the
arithmetic done by hand-written CUDA kernels behind the C ABI in `include/cfa_b200.h`.
Attributes resolve lazily so that `python -m clip_finegrained_alignment_b200.build` can run before
the shared library exists; touching any compute class without the library raises ImportError.
"""
__all__ = ["SPARCLoss", "CustomCLIPLoss", "CLIPCountLoss", "CountLoss", "AdamSPD", "CLIPFineTuneConfig"]


def __getattr__(name):
    if name in ("SPARCLoss", "CustomCLIPLoss", "CLIPCountLoss", "CountLoss"):
        from . import losses
        return getattr(losses, name)
    if name == "AdamSPD":
        from .optimizers import AdamSPD
        return AdamSPD
    if name == "CLIPFineTuneConfig":
        from .config import CLIPFineTuneConfig
        return CLIPFineTuneConfig
    raise AttributeError(name)
