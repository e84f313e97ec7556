"""B200-native (sm_100a) Faster R-CNN region stage: libfrr.so kernels behind the reference's call sites.

Importing the package does not need a GPU; calling any op does, and raises if libfrr.so is missing."""
__all__ = ["ops", "region", "targets", "modules", "anchor", "synth", "dist"]
