"""Import the upstream reference (read-only at /root/reference) in THIS container.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py and oracle
validation scripts to run the reference's own PyTorch code on CPU.  The
reference is not present on the GPU box, so nothing at run time
(`-m gpu` tests, smoke(), bench.py) may import this module.

Recipe: the reference cannot be imported as shipped --
`models/text_encoder.py:4` does `from turtle import forward` (needs tkinter)
and `models/__init__.py:6,9` imports two modules whose sources are missing.
Three stub modules injected into sys.modules make `import models` succeed.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ACVAE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def load_reference():
    """Return (models, train_util) modules of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if "turtle" not in sys.modules:
        turtle = types.ModuleType("turtle")
        turtle.forward = lambda *a, **k: None
        sys.modules["turtle"] = turtle
    for name in ("models.transformer_model", "models.transformer_vae_model"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    import models  # noqa: E402  (the reference's package)
    import utils.train_util as train_util  # noqa: E402
    return models, train_util
