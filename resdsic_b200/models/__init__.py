"""Model registry -- mirror of the reference's `compress/models/__init__.py:22-62`
for the model families BASELINE.json names (`-m/-a cnn|stf`)."""
from .wacnn import WACNN

models = {
    "cnn": WACNN,
}


def configure_model(args):
    """reference models/__init__.py:33-62: `models[args.model](N=args.N, M=args.M)`."""
    name = args.model
    if name not in models:
        raise KeyError(f"unknown model {name!r}; available: {sorted(models)}")
    return models[name](N=getattr(args, "N", 192), M=getattr(args, "M", 320))


__all__ = ["models", "configure_model", "WACNN"]
