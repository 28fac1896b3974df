"""Hot-path operators of the B200 engine: every op here is a hand-written sm_100a kernel (``internevo_b200/csrc``)
with a plain PyTorch reference used only on CPU (host-logic tests) and as the numerics oracle."""
from . import _lib  # noqa: F401
from .adam import adamw_, clip_scalars_, sumsq_  # noqa: F401
from .cross_entropy import cross_entropy  # noqa: F401
from .gemm import launch_count, linear, linear_gelu, matmul, matmul_gelu, matmul_swiglu  # noqa: F401
from .norm import LayerNorm, RMSNorm, add_rmsnorm, dropout_add_layernorm, layernorm_ref, rmsnorm_ref  # noqa: F401
from .rope import RotaryTables, apply_rotary_packed, rope_  # noqa: F401
from .grouped import aligned_offsets, grouped_matmul, grouped_swiglu_mlp, grouped_wgrad  # noqa: F401
from .swiglu import silu_mul, swiglu_interleaved  # noqa: F401
