"""Soft dependency on `lightning` (reference: `from lightning import LightningModule`, genie/tokenizer.py:13).
When the package is installed the public classes subclass the real LightningModule so a Lightning Trainer
can drive them unchanged; otherwise a minimal stand-in keeps the same method surface."""
import torch.nn as nn

try:  # pragma: no cover - depends on the environment
    from lightning import LightningModule  # type: ignore
    HAVE_LIGHTNING = True
except Exception:  # lightning is not installed in the build image
    HAVE_LIGHTNING = False

    class LightningModule(nn.Module):
        """Stand-in with the hooks the reference modules call (save_hyperparameters / log_dict / log)."""

        def __init__(self, *a, **k):
            super().__init__()
            self.logged = {}

        def save_hyperparameters(self, *a, **k):
            pass

        def log_dict(self, d, *a, **k):
            self.logged.update({k_: (v.detach() if hasattr(v, 'detach') else v) for k_, v in d.items()})

        def log(self, name, value, *a, **k):
            self.logged[name] = value.detach() if hasattr(value, 'detach') else value
