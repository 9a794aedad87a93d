"""rdb200 -- host side of the B200-native Reflected-Diffusion sampling path.

`_lib`      ctypes binding of librdb200.so (the C ABI in include/rdb200.h)
`ops`       tensor-level wrappers of the element-wise kernels
`pack`      reference state_dict -> kernel weight layouts
`planner`   NCSN++ topology -> flat op plan (mirrors NCSNpp.__init__ / forward)
`engine`    plan + workspace + sampler graph for one (model, batch, H, W)
`dist`      one-process-per-GPU batch sharding + final NCCL all-gather
"""
from . import _lib  # noqa: F401
