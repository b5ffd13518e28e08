"""dcnv3-b200: the DCNv3 deformable-sampling core of Z1HaoC/YOLO-Dual, rebuilt for B200.

Importable as ``yolo_dual_b200`` (a hyphen cannot appear in a Python package name;
``yolo-dual_b200`` in the repo root is a symlink to this directory).

    from yolo_dual_b200.ops_dcnv3.modules import DCNv3
    from yolo_dual_b200.ops_dcnv3.functions import DCNv3Function
"""
__version__ = "0.1.0"
