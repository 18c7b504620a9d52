"""cerebralsignalnetworks_b200 -- the EEG distillation train step of Vi-Sri/CerebralSignalNetworks on B200.

Host side mirrors the reference's module API (models.lstm.Model, DINOHead, DINOLoss, EEGFilters); the arithmetic
lives in libcsn_b200.so (hand-written sm_100a CUDA behind the C-ABI of include/csn_b200.h).  There is no CPU
fallback: importing works anywhere, calling needs the built library and a B200.
"""
from . import _lib
from ._lib import CsnError, LIB_PATH

__all__ = ["CsnError", "LIB_PATH", "Model", "DINOHead", "DINOLoss", "MultiCropWrapper", "EEGFilters",
           "DistillTrainStep", "MultiCropDistillStep", "ops", "IndexFlatL2", "IndexFlatIP", "retrieval",
           "FeatureDistributionLoss", "CosineSimilarityLoss", "HyperParams", "loss_fn_kd", "loss_fn_kd_hinton",
           "cosine_similarity_loss"]


def __getattr__(name):  # lazy: keep `import cerebralsignalnetworks_b200` cheap and torch-free until used
    if name in ("Model", "LSTMStack", "Linear"):
        from . import lstm
        return getattr(lstm, name)
    if name in ("DINOHead", "DINOLoss", "MultiCropWrapper"):
        from . import dino
        return getattr(dino, name)
    if name in ("EEGFilters", "butter_bandpass_sos"):
        from . import eeg_filters
        return getattr(eeg_filters, name)
    if name == "DistillTrainStep":
        from .train_step import DistillTrainStep
        return DistillTrainStep
    if name == "MultiCropDistillStep":
        from .multicrop_step import MultiCropDistillStep
        return MultiCropDistillStep
    if name in ("FeatureDistributionLoss", "CosineSimilarityLoss", "HyperParams", "loss_fn_kd", "loss_fn_kd_hinton",
                "cosine_similarity_loss"):
        from . import losses
        return getattr(losses, name)
    if name in ("IndexFlatL2", "IndexFlatIP", "retrieval"):
        import importlib
        mod = importlib.import_module(".retrieval", __name__)
        return mod if name == "retrieval" else getattr(mod, name)
    if name == "ops":
        import importlib
        return importlib.import_module(".ops", __name__)
    raise AttributeError(name)
