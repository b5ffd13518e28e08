"""Generate the golden fixtures from the REFERENCE's own dcnv3_core_pytorch.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):   python tests/golden/make_golden.py

The reference file does ``import DCNv3`` (its missing pybind module) at import
time (dcnv3_func.py:16); a stub module is installed first.  Nothing is written
outside tests/golden/.
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

REF_LIB = "/root/reference/models/ops_dcnv3/build/lib.linux-x86_64-cpython-38"


def load_reference():
    sys.modules.setdefault("DCNv3", types.ModuleType("DCNv3"))
    sys.path.insert(0, REF_LIB)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from functions.dcnv3_func import dcnv3_core_pytorch  # the reference itself
    return dcnv3_core_pytorch


def main():
    from cases import CASES, op_args
    from oracle.dcnv3_oracle import make_inputs

    ref = load_reference()
    torch.set_num_threads(1)  # deterministic reductions
    total = 0
    for c in CASES:
        dt = getattr(torch, c["dtype"])
        # inputs are drawn in fp32 and cast, as test.py does (.double() at :42-44)
        x, off, m, go = make_inputs(c["N"], c["H"], c["W"], c["G"], c["gc"], c["kh"], c["kw"],
                                    c["sh"], c["sw"], c["ph"], c["pw"], c["dh"], c["dw"],
                                    dist=c["dist"], seed=c["seed"], dtype=torch.float32)
        x, off, m, go = (t.to(dt) for t in (x, off, m, go))
        xi, oi, mi = (t.clone().requires_grad_(True) for t in (x, off, m))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = ref(xi, oi, mi, *op_args(c))
            out.backward(go)
        path = os.path.join(HERE, c["name"] + ".npz")
        np.savez_compressed(
            path, input=x.numpy(), offset=off.numpy(), mask=m.numpy(), grad_out=go.numpy(),
            output=out.detach().numpy(), grad_input=xi.grad.numpy(),
            grad_offset=oi.grad.numpy(), grad_mask=mi.grad.numpy())
        sz = os.path.getsize(path)
        total += sz
        print(f"{c['name']:24s} out{tuple(out.shape)} {sz/1024:.1f} KiB")
    print(f"total {total/1024:.1f} KiB; torch {torch.__version__}")


if __name__ == "__main__":
    main()
