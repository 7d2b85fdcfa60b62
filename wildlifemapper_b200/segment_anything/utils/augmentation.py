"""Training-time augmentation lives outside the tile-detection hot path.  The module exists so
that the reference data loader's imports resolve (dataloader_coco.py:19-20); its functions raise when called."""


def _out_of_scope(*_a, **_k):
    raise NotImplementedError("training augmentation is out of scope of the B200 inference hot path")


random_perspective = _out_of_scope
Compose = ToTensor = Normalize = RandomHorizontalFlip = RandomResize = _out_of_scope
