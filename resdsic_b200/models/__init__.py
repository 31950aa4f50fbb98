"""Model registry -- mirror of the reference's `compress/models/__init__.py:22-62`
for the model families BASELINE.json names (`-m/-a cnn|stf`)."""
from .stf import SymmetricalTransFormer
from .wacnn import WACNN

models = {
    "cnn": WACNN,
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
    return models[name](N=getattr(args, "N", 192), M=getattr(args, "M", 320))


__all__ = ["models", "configure_model", "WACNN", "SymmetricalTransFormer"]
