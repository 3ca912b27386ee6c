"""mimrl_b200 — B200-native MI / CMI hot path of kiva12138/MIMRL.

Drop-in modules (same names and signatures as the reference's Python symbols):

    from mimrl_b200.vmi import CriticModel, BaselineModel, infonce_lower_bound, ...   # VMI.py
    from mimrl_b200.model import VMIEstimator, VCMIEstimator, prod_knn_sample, ...    # Model.py:47-225
    from mimrl_b200.mlp_process import MLP, MLPsBlock, MLPEncoder                     # MLPProcess.py

All arithmetic of the path runs in hand-written sm_100a kernels in
``lib/libmimrl_b200.so`` (``python -m mimrl_b200.build``); there is no CPU or
eager-PyTorch fallback.  Submodules are imported lazily so that the build
script can be run before the library exists.
"""
import importlib

__all__ = ["vmi", "model", "mlp_process", "rowblock", "build", "_lib"]


def __getattr__(name):
    if name in __all__:
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
