"""sam2_unet_b200 — B200-native (sm_100a) SAM2-UNet training / inference hot path.

Public surface (mirrors /root/reference/SAM2UNet.py and train.py):
    SAM2UNet(checkpoint_path="", *, model_cfg="sam2_hiera_s.yaml", dtype="bf16")   nn.Module, forward -> out, out1, out2
    structure_loss(pred, mask)                                                     train.py:21-29
    FusedAdamW, TrainStep                                                          train.py:48-52,66-86
    TrainAugment, preprocess_image                                                 dataset.py:288-313, :336-407
    infer_tail                                                                     test.py:66-76
    evaluate_segmentation_performance, evaluate_dataset, print_eval_report         eval.py:23-224
"""
from . import _lib  # noqa: F401
from .config import trunk_config  # noqa: F401
from .loss import structure_loss, structure_loss3  # noqa: F401
from .model import SAM2UNet  # noqa: F401
from .optim import FusedAdamW, Predictor, TrainStep, cosine_lr  # noqa: F401
from .evalmetrics import evaluate_dataset, evaluate_segmentation_performance, print_eval_report  # noqa: F401
from .postprocess import infer_tail, preprocess_image  # noqa: F401
from .augment import TrainAugment, draw_train_params  # noqa: F401

__all__ = ["SAM2UNet", "structure_loss", "structure_loss3", "FusedAdamW", "TrainStep", "Predictor", "cosine_lr",
           "infer_tail", "preprocess_image", "TrainAugment", "draw_train_params", "evaluate_segmentation_performance",
           "evaluate_dataset", "print_eval_report", "trunk_config"]
