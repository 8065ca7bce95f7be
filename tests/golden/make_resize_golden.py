"""Golden vectors for the input-side resize (oracle/resize_oracle.py, csrc/resize.cu): outputs of Pillow's
Image.resize((100, 32), BILINEAR) - what torchvision.transforms.Resize((32, 100)) runs on the reference's PIL
crops (coco_dataset.py:27,40) - for seeded random crops of assorted sizes, plus torchvision's ToTensor of one
of them.  Run in the build container:  python tests/golden/make_resize_golden.py"""
import os

import numpy as np
import PIL
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
SHAPES = [(57, 233), (32, 100), (32, 250), (90, 100), (7, 13), (1, 1), (5, 400), (300, 40), (33, 101), (31, 99),
          (64, 256), (200, 1000), (17, 2000), (1000, 30), (2, 3), (16, 50), (45, 45), (24, 75), (128, 512)]


def make_crop(seed, h, w):
    """Seeded text-like crop: low-frequency field + noise, with exact 0 and 255 pixels."""
    rs = np.random.RandomState(seed)
    gh, gw = max(2, h // 8 + 1), max(2, w // 8 + 1)
    coarse = rs.rand(gh, gw)
    yi = np.minimum((np.arange(h) * gh) // max(h, 1), gh - 1)
    xi = np.minimum((np.arange(w) * gw) // max(w, 1), gw - 1)
    a = 0.8 * coarse[yi][:, xi] + 0.2 * rs.rand(h, w)
    a = np.clip(a * 255.0, 0, 255).astype(np.uint8)
    if h * w > 4:
        a[rs.randint(h), rs.randint(w)] = 255
        a[rs.randint(h), rs.randint(w)] = 0
    return a


if __name__ == "__main__":
    out = {"shapes": np.array(SHAPES, np.int32), "pillow_version": np.array(PIL.__version__)}
    for i, (h, w) in enumerate(SHAPES):
        a = make_crop(1000 + i, h, w)
        out["resized_%d" % i] = np.asarray(Image.fromarray(a, "L").resize((100, 32), Image.BILINEAR))
    from torchvision import transforms
    a = make_crop(1000, *SHAPES[0])
    t = transforms.ToTensor()(transforms.Resize((32, 100))(Image.fromarray(a, "L")))
    out["to_tensor_0"] = t.numpy()
    np.savez_compressed(os.path.join(HERE, "resize_cases.npz"), **out)
    print("wrote resize_cases.npz", os.path.getsize(os.path.join(HERE, "resize_cases.npz")) // 1024, "KiB")
