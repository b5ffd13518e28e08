"""TEST INFRASTRUCTURE.  Recipe that compiles the reference's CUDA DCNv3 kernels for sm_100a into
oracle/_ref/dcnv3_ref_cuda.so (see oracle/ref_cuda_shim.cu).  Runs only where /root/reference is
mounted (the build container); the GPU box uses the prebuilt file.

    python oracle/build_ref_cuda.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/models/ops_dcnv3/src"
OUT = os.path.join(HERE, "_ref")
NAME = "dcnv3_ref_cuda"


def so_path():
    return os.path.join(OUT, NAME + ".so")


def build(verbose=False):
    if not os.path.isdir(REF_SRC):
        return None
    if os.path.exists(so_path()) and os.path.getmtime(so_path()) > os.path.getmtime(os.path.join(HERE, "ref_cuda_shim.cu")):
        return so_path()
    os.makedirs(OUT, exist_ok=True)
    os.environ.pop("CC", None)   # the image's gcc wrapper breaks nvcc's host compile
    os.environ.pop("CXX", None)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load
    load(name=NAME, sources=[os.path.join(HERE, "ref_cuda_shim.cu")], extra_include_paths=[REF_SRC],
         extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a", "-O3"],
         build_directory=OUT, verbose=verbose, is_python_module=False)
    return so_path() if os.path.exists(so_path()) else None


def load_module():
    """Import the prebuilt extension (GPU box or here); None when it was never built."""
    p = so_path()
    if not os.path.exists(p):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="--verbose" in sys.argv))
