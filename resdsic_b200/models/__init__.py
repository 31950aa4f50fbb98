"""Model registry -- mirror of the reference's `compress/models/__init__.py:22-62`
for the model families BASELINE.json names (`-m/-a cnn|stf`)."""
from .scalable import (ResWACNNIndependentEntropy, conditional_scalable_icd, conditional_scalable_imd, scalable_icd,
                       scalable_imd)
from .stf import SymmetricalTransFormer
from .wacnn import WACNN

models = {
    "cnn": WACNN,
    # ResDSIC scalable codecs: every key of the reference registry (models/__init__.py:22-29)
    "icd": scalable_icd,
    "imd": scalable_imd,
    "cicd": conditional_scalable_icd,
    "cimd": conditional_scalable_imd,
    "ind": ResWACNNIndependentEntropy,
    # the reference README's `-m stf`; the reference tree itself has no STF model, so this
    # entry is builder-defined (see models/stf.py) -- every block it is made of is pinned to the reference
    "stf": SymmetricalTransFormer,
}


def configure_model(args):
    """reference models/__init__.py:33-62: `models[args.model](N=args.N, M=args.M)`."""
    name = args.model
    if name not in models:
        raise KeyError(f"unknown model {name!r}; available: {sorted(models)}")
    if name == "stf":
        return models[name]()  # STF fixes its widths (N=192, M=384), like the upstream `-m stf`
    if name == "ind":  # reference models/__init__.py:35-43
        return models[name](N=getattr(args, "N", 192), M=getattr(args, "M", 320),
                            mask_policy=getattr(args, "mask_policy", "two-levels"),
                            lambda_list=getattr(args, "lambda_list", [0.0035, 0.065]),
                            lrp_prog=getattr(args, "lrp_prog", True),
                            independent_lrp=getattr(args, "independent_lrp", False),
                            multiple_decoder=getattr(args, "multiple_decoder", False))
    if name in ("cicd", "cimd"):  # reference models/__init__.py:45-51
        return models[name](N=getattr(args, "N", 192), M=getattr(args, "M", 320),
                            mask_policy=getattr(args, "mask_policy", "two-levels"),
                            lambda_list=getattr(args, "lambda_list", [0.0035, 0.065]),
                            joiner_policy=getattr(args, "joiner_policy", "conditional"))
    if name in ("icd", "imd"):  # reference models/__init__.py:49-54
        return models[name](N=getattr(args, "N", 192), M=getattr(args, "M", 320),
                            mask_policy=getattr(args, "mask_policy", "two-levels"),
                            lambda_list=getattr(args, "lambda_list", [0.0035, 0.065]))
    return models[name](N=getattr(args, "N", 192), M=getattr(args, "M", 320))


__all__ = ["models", "configure_model", "WACNN", "SymmetricalTransFormer", "scalable_icd", "scalable_imd", "conditional_scalable_icd",
           "conditional_scalable_imd", "ResWACNNIndependentEntropy"]
