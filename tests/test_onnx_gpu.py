"""ONNX export of the DCNv3 op: the reference's `DCNv3Function.symbolic` emits one `mmdeploy::TRTDCNv3` node with twelve
attributes (LIB/functions/dcnv3_func.py:63-89).  The `onnx` package is not in this image, so the test drives torch's
TorchScript exporter up to the ONNX graph (`_model_to_graph`: trace -> symbolic functions -> torch._C.Graph) and
inspects the node instead of serialising a file.  Tracing executes the forward, hence a GPU test."""
import warnings

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _onnx_graph(module, args):
    from torch.onnx._internal.torchscript_exporter import utils as U
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        graph, _params, _out = U._model_to_graph(module, args)
    return graph


class _Core(torch.nn.Module):
    def __init__(self, fn, extra):
        super().__init__()
        self.fn, self.extra = fn, extra

    def forward(self, x, off, m):
        return self.fn.apply(x, off, m, *self.extra)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16], ids=["f32", "f16"])
def test_symbolic_emits_the_reference_node(dtype):
    from yolo_dual_b200.ops_dcnv3.functions import DCNv3Function
    N, H, W, G, gc = 1, 6, 7, 4, 16
    extra = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.5, 256)
    x = torch.randn(N, H, W, G * gc, device=DEV, dtype=dtype)
    off = torch.randn(N, H, W, G * 9 * 2, device=DEV, dtype=dtype)
    m = torch.softmax(torch.randn(N, H, W, G, 9, device=DEV), -1).reshape(N, H, W, G * 9).to(dtype)
    graph = _onnx_graph(_Core(DCNv3Function, extra), (x, off, m))
    nodes = [n for n in graph.nodes() if n.kind() == "mmdeploy::TRTDCNv3"]
    assert len(nodes) == 1, [n.kind() for n in graph.nodes()]
    n = nodes[0]
    assert len(list(n.inputs())) == 3 and len(list(n.outputs())) == 1
    want_i = dict(kernel_h=3, kernel_w=3, stride_h=1, stride_w=1, pad_h=1, pad_w=1, dilation_h=1, dilation_w=1,
                  group=G, group_channels=gc, im2col_step=256)
    assert sorted(n.attributeNames()) == sorted(list(want_i) + ["offset_scale"])
    for k, v in want_i.items():
        assert n.i(k) == v, k
    assert abs(n.f("offset_scale") - 1.5) < 1e-7


def test_module_exports_one_node_per_site():
    """The nn.Module as a whole: Linear / depthwise-conv glue becomes stock ONNX ops, the core one custom node."""
    from yolo_dual_b200.ops_dcnv3.modules import DCNv3
    mod = DCNv3(channels=64, group=4).to(DEV).eval()
    x = torch.randn(1, 8, 8, 64, device=DEV)
    graph = _onnx_graph(mod, (x,))
    kinds = [n.kind() for n in graph.nodes()]
    assert kinds.count("mmdeploy::TRTDCNv3") == 1
    assert any(k == "onnx::Softmax" for k in kinds) and any(k in ("onnx::Gemm", "onnx::MatMul") for k in kinds)
