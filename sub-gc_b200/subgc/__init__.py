"""subgc — B200-native implementation of the Sub-GC captioning hot path (GCN sub-graph encoder, sGPN scorer,
top-down attention-LSTM decoder) behind the reference's `models.setup(opt)` / `AttModel.forward(mode=...)` API.

Importing the package is cheap (no CUDA needed); the C-ABI library `libsubgc_b200.so` is loaded on first use by
`subgc._lib` and every compute entry point fails loudly if it is missing — there is no CPU fallback.
"""
from .config import Dims, SMALL, dims_from_opt, make_opt  # noqa: F401

__version__ = "0.1.0"
