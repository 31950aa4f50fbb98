"""Rate-distortion training step of `-m cnn` (BASELINE config 4) on the B200 kernel library, fp32.

    net = resdsic_b200.WACNN().cuda()
    out = resdsic_b200.training.train_forward(net, x)          # reference cnn.py:143-193 in train() mode
    crit = resdsic_b200.training.RateDistortionLoss(lmbda=0.0035)
    crit(out, x)["loss"].backward()                            # every node: a CUDA kernel of this library
    resdsic_b200.training.aux_loss(net).backward()

(reference training/step.py:42-56).  Multi-GPU: one process per GPU + `GradBucketReducer` (ddp.py).
"""
from .ddp import GradBucketReducer
from .loss import RateDistortionLoss, ScalableRateDistortionLoss
from .model import aux_loss, train_forward

__all__ = ["train_forward", "aux_loss", "RateDistortionLoss", "ScalableRateDistortionLoss", "GradBucketReducer"]
