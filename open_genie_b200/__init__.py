"""open_genie_b200 — B200-native (sm_100a) implementation of open-genie's data-parallel hot path behind the
reference's own Python surface (myscience/open-genie: genie/__init__.py).

Importing this package never touches the GPU; the CUDA library (csrc/libopengenie_b200.so) is loaded on
first use and there is no CPU or PyTorch fallback behind it."""
from .tokenizer import MAGVIT2_DEC_DESC, MAGVIT2_ENC_DESC, REPR_TOK_DEC, REPR_TOK_ENC, VideoTokenizer
from .module import get_module, parse_blueprint
from .optim import FusedAdamW

__all__ = ['VideoTokenizer', 'MAGVIT2_ENC_DESC', 'MAGVIT2_DEC_DESC', 'REPR_TOK_ENC', 'REPR_TOK_DEC',
           'get_module', 'parse_blueprint', 'FusedAdamW']
