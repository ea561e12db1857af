"""B200-native multi-stream tracking-by-detection hot path - a drop-in for BoxMOT's
``create_tracker(...)`` / ``tracker.update(dets, img) -> ndarray[M, 8]`` API.  See DESIGN.md.

Importing the package does not touch the GPU; the CUDA library is loaded on first use and
there is no CPU fallback.
"""
__version__ = "0.1.0"

from .tracker_zoo import create_tracker, get_tracker_config  # noqa: E402,F401

TRACKERS = ["bytetrack", "botsort", "ocsort", "strongsort", "deepocsort", "hybridsort"]


def __getattr__(name):          # lazy: BYTETracker / OCSORT / BoTSORT / BatchedTracker
    if name == "BYTETracker":
        from .trackers.bytetrack import BYTETracker
        return BYTETracker
    if name in ("OCSORT", "OCSort"):
        from .trackers.ocsort import OCSort
        return OCSort
    if name == "BoTSORT":
        from .trackers.botsort import BoTSORT
        return BoTSORT
    if name == "StrongSORT":
        from .trackers.strongsort import StrongSORT
        return StrongSORT
    if name in ("DeepOCSORT", "DeepOCSort"):
        from .trackers.deepocsort import DeepOCSort
        return DeepOCSort
    if name == "HybridSORT":
        from .trackers.hybridsort import HybridSORT
        return HybridSORT
    if name == "BatchedTracker":
        from .batch import BatchedTracker
        return BatchedTracker
    raise AttributeError(name)


__all__ = ("__version__", "BYTETracker", "OCSORT", "BoTSORT", "StrongSORT", "DeepOCSORT", "HybridSORT", "BatchedTracker", "create_tracker",
           "get_tracker_config", "TRACKERS")
