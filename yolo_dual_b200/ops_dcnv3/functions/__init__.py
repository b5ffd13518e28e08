from .dcnv3_func import (DCNv3Function, DCNv3PackedFunction, DCNv3SoftmaxFunction, dcnv3_debug_indices,
                         get_grad_accum, set_grad_accum)

__all__ = ["DCNv3Function", "DCNv3SoftmaxFunction", "DCNv3PackedFunction", "dcnv3_debug_indices", "set_grad_accum",
           "get_grad_accum"]
