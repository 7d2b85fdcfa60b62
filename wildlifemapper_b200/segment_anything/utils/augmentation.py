"""Host-side loader transforms with the reference's names and call contract (reference ``utils/augmentation.py``), limited to
what its data loader composes (``dataloader_coco.py:275-292``): ``Compose``, ``RandomResize`` / ``resize``, ``ToTensor``,
``Normalize``, ``FlipLR``, ``RandomHorizontalFlip`` / ``hflip``.  They run in the DataLoader workers on PIL images and CPU
tensors, like the reference's -- the tile-detection hot path starts behind them.  (The same resize + normalise + pad is
available on the device for whole survey images: ``wildlifemapper_b200.survey``.)  The crop / erase / select / pad
augmentations of the DETR transform file are not part of the reference's pipelines and raise when used."""
import random

import torch
import torchvision.transforms.functional as F


def _target_size(image_size, size, max_size=None):
    """(height, width) of the resized image: the short side becomes ``size`` unless that would push the long side over
    ``max_size`` (then the long side becomes ``max_size``); ``size`` given as (w, h) is taken as is."""
    if isinstance(size, (list, tuple)):
        return size[::-1]
    w, h = image_size
    if max_size is not None:
        short, long_ = float(min(w, h)), float(max(w, h))
        if long_ / short * size > max_size:
            size = int(round(max_size * short / long_))
    if (w <= h and w == size) or (h <= w and h == size):
        return (h, w)
    if w < h:
        return (int(size * h / w), size)
    return (size, int(size * w / h))


def resize(image, target, size, max_size=None):
    """PIL image (+ target dict) -> resized image, target with boxes / area scaled and ``size`` = (h, w) of the result."""
    out_hw = _target_size(image.size, size, max_size)
    resized = F.resize(image, out_hw)
    if target is None:
        return resized, None
    rw, rh = (float(a) / float(b) for a, b in zip(resized.size, image.size))
    target = target.copy()
    if "boxes" in target:
        target["boxes"] = target["boxes"] * torch.as_tensor([rw, rh, rw, rh])
    if "area" in target:
        target["area"] = target["area"] * (rw * rh)
    h, w = out_hw
    target["size"] = torch.tensor([h, w])
    return resized, target


def hflip(image, target):
    flipped = F.hflip(image)
    w, _h = image.size
    target = target.copy()
    if "boxes" in target:
        b = target["boxes"]
        target["boxes"] = b[:, [2, 1, 0, 3]] * torch.as_tensor([-1, 1, -1, 1]) + torch.as_tensor([w, 0, w, 0])
    if "masks" in target:
        target["masks"] = target["masks"].flip(-1)
    return flipped, target


class RandomResize(object):
    def __init__(self, sizes, max_size=None):
        assert isinstance(sizes, (list, tuple))
        self.sizes = sizes
        self.max_size = max_size

    def __call__(self, img, target=None):
        return resize(img, target, random.choice(self.sizes), self.max_size)


class RandomHorizontalFlip(object):
    def __init__(self, p=0.5):
        self.p = p

    def __call__(self, img, target):
        if random.random() < self.p:
            return hflip(img, target)
        return img, target


class ToTensor(object):
    def __call__(self, img, target):
        return F.to_tensor(img), target


class Normalize(object):
    """Channel normalisation of the image tensor; boxes go from absolute xyxy to (cx, cy, w, h) relative to the image size,
    ``center`` points to relative coordinates."""

    def __init__(self, mean, std):
        self.mean = mean
        self.std = std

    def __call__(self, image, target=None):
        image = F.normalize(image, mean=self.mean, std=self.std)
        if target is None:
            return image, None
        target = target.copy()
        h, w = image.shape[-2:]
        if "center" in target:
            target["center"] = target["center"] / torch.tensor([w, h], dtype=torch.float32)
        if "boxes" in target:
            x0, y0, x1, y1 = target["boxes"].unbind(-1)
            cxcywh = torch.stack([(x0 + x1) / 2, (y0 + y1) / 2, x1 - x0, y1 - y0], dim=-1)
            target["boxes"] = cxcywh / torch.tensor([w, h, w, h], dtype=torch.float32)
        return image, target


class FlipLR(object):
    """The reference's train-time flip: ``torch.fliplr`` on the [C, H, W] tensor, i.e. along H, with cy -> 1 - cy."""

    def __init__(self, fliplr=0.5):
        self.fliplr = fliplr

    def __call__(self, image, target=None):
        if random.random() < self.fliplr:
            image = torch.fliplr(image)
            if target is None:
                return image, None
            for key in ("boxes", "center"):
                if key in target and len(target[key]):
                    target[key][:, 1] = 1 - target[key][:, 1]
        return image, target


class Compose(object):
    def __init__(self, transforms):
        self.transforms = transforms

    def __call__(self, image, target):
        for t in self.transforms:
            image, target = t(image, target)
        return image, target

    def __repr__(self):
        return self.__class__.__name__ + "(" + "".join("\n    {0}".format(t) for t in self.transforms) + "\n)"


def _not_in_the_reference_pipelines(*_a, **_k):
    raise NotImplementedError("this DETR augmentation is not used by the reference's loader pipelines "
                              "(dataloader_coco.py:275-292) and is not provided")


crop = pad = RandomCrop = RandomSizeCrop = CenterCrop = RandomPad = RandomSelect = RandomErasing = _not_in_the_reference_pipelines
