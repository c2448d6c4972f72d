"""Synthetic calibration inputs: there is no dataset on the GPU box.

images  torch.randn(3, H, W) fp32 seeded 1234 + global image index (normalised-image-like,
        datasets/Base.py:91-96 normalises real images to zero mean / unit std)
labels  [H, W] uint8 in [0, K) U {255}: spatially coherent Voronoi blobs whose classes follow a
        long-tail (Zipf) prior, plus ~3 % ignore pixels in a few rectangles.  i.i.d. labels would
        destroy the run-length structure real segmentation maps have.
Everything depends only on (global image index, K, H, W), never on the rank count.
"""
import numpy as np
import torch

IGNORE_LABEL = 255


def class_prior(num_classes):
    """Zipf prior; the exponent makes max/min ~ 370 at 19 classes like Cityscapes' pixel totals
    (datasets/CSdatasets.py:23-27) and keeps a long tail at 150 / 171 classes."""
    s = 2.0 if num_classes <= 32 else 1.2
    p = 1.0 / np.arange(1, num_classes + 1, dtype=np.float64) ** s
    return p / p.sum()


def synthetic_labels(index, num_classes, height, width, n_blobs=48, ignore_frac=0.03):
    rng = np.random.RandomState(977 + 31 * int(index))
    ys = rng.randint(0, height, size=n_blobs)
    xs = rng.randint(0, width, size=n_blobs)
    cls = rng.choice(num_classes, size=n_blobs, p=class_prior(num_classes))
    # anisotropic metric -> elongated regions (roads, sky bands)
    ay = rng.uniform(0.5, 2.0, size=n_blobs).astype(np.float32)
    ax = rng.uniform(0.5, 2.0, size=n_blobs).astype(np.float32)
    yy = np.arange(height, dtype=np.float32)[:, None, None]
    xx = np.arange(width, dtype=np.float32)[None, :, None]
    best = np.full((height, width), np.inf, dtype=np.float32)
    lab = np.zeros((height, width), dtype=np.uint8)
    for b in range(n_blobs):  # streaming arg-min keeps memory at O(H*W)
        d = ay[b] * (yy[:, :, 0] - ys[b]) ** 2 + ax[b] * (xx[:, :, 0] - xs[b]) ** 2
        m = d < best
        best[m] = d[m]
        lab[m] = cls[b]
    target = ignore_frac * height * width
    done = 0
    while done < target:
        rh = rng.randint(max(height // 32, 1), max(height // 8, 2))
        rw = rng.randint(max(width // 32, 1), max(width // 8, 2))
        y0 = rng.randint(0, height - rh + 1)
        x0 = rng.randint(0, width - rw + 1)
        lab[y0:y0 + rh, x0:x0 + rw] = IGNORE_LABEL
        done += rh * rw
    return torch.from_numpy(lab)


def synthetic_images(index, height, width):
    g = torch.Generator().manual_seed(1234 + int(index))
    return torch.randn(3, height, width, generator=g)


def synthetic_batch(indices, num_classes, height, width, label_dtype=torch.uint8):
    imgs = torch.stack([synthetic_images(i, height, width) for i in indices])
    labs = torch.stack([synthetic_labels(i, num_classes, height, width) for i in indices]).to(label_dtype)
    return imgs, labs
