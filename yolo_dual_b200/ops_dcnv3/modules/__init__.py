from .dcnv3 import DCNv3

__all__ = ["DCNv3"]
