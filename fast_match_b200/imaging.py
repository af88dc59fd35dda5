"""Image loading / scaling helpers (host glue, out of the hot path).

Mirrors the reference's imaging.py:17-85 (scale classes, get_thumbnail, get_size,
open_img) on Python 3 / PIL >= 10 / cv2 4.
"""
import cv2
import numpy
from PIL import Image


def _target_size(width, height, size):
    """Aspect-preserving size with the long side pinned (imaging.py:29-36)."""
    if width > height:
        w = size[0]
        h = int((w / float(width)) * height)
    else:
        h = size[1]
        w = int((h / float(height)) * width)
    return (int(w), int(h))


def _as_pil(data):
    if isinstance(data, str):
        return Image.open(data)
    return Image.fromarray(data)


def get_thumbnail(path_or_array, size=(200, 200)):
    """PIL two-pass antialiased thumbnail (imaging.py:49-55, 67-69)."""
    img = _as_pil(path_or_array)
    new_size = _target_size(img.size[0], img.size[1], size)
    img.thumbnail(tuple(2 * i for i in new_size))
    img.thumbnail(new_size, Image.LANCZOS)  # Image.ANTIALIAS was an alias of LANCZOS
    return numpy.array(img, dtype=numpy.uint8)


def get_size(path_or_array):
    """(width, height) (imaging.py:71-72)."""
    return _as_pil(path_or_array).size


def open_img(path, size=None):
    """cv2.imread, optionally area-resized so the long side is `size` (imaging.py:80-85)."""
    img = cv2.imread(path)
    if img is None:
        raise IOError("cannot read image %r" % (path,))
    if size is None or size == -1:
        return img
    if not isinstance(size, (tuple, list)):
        size = (size, size)
    new_size = _target_size(img.shape[1], img.shape[0], size)
    return cv2.resize(img, new_size, interpolation=cv2.INTER_AREA)
