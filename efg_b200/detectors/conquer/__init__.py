from .model import ConQueR, build_model

__all__ = ["ConQueR", "build_model"]
