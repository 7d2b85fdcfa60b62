"""``SamPredictor`` must stay importable (reference ``segment_anything/__init__.py:14``).  In the reference it is an
upstream-SAM leftover that calls the modified modules with the old signatures and raises;
it is outside the tile-detection hot path."""


class SamPredictor:
    def __init__(self, *args, **kwargs) -> None:
        raise NotImplementedError(
            "SamPredictor is an unused upstream-SAM leftover in the reference (predictor.py:89,229 call the modified "
            "modules with stale signatures); it is out of scope for the tile-detection hot path. Use MedSAM.forward.")
