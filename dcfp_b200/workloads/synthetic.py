"""Synthetic calibration inputs: there is no dataset on the GPU box.

images  torch.randn(3, H, W) fp32 seeded 1234 + global image index (normalised-image-like,
        datasets/Base.py:91-96 normalises real images to zero mean / unit std)
labels  [H, W] uint8 in [0, K) U {255}: spatially coherent Voronoi blobs whose classes follow a
        long-tail (Zipf) prior, plus ~3 % ignore pixels in a few rectangles.  i.i.d. labels would
        destroy the run-length structure real segmentation maps have.
        `fragmentation` picks how fine the regions are (FRAGMENTATION): "coarse" = 48 blobs (the round-1 workload: at
        1/8 resolution a class run is hundreds of pixels long), "street" = 400 blobs + 60 thin structures (poles, lane
        markings, wires: what a Cityscapes frame looks like -- bench.py's default), "fine" = 3000 blobs + 300 thin
        structures (a class boundary every few pixels at 1/8 resolution: stress for K1's slot cache).
Everything depends only on (global image index, K, H, W, fragmentation), never on the rank count.
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


FRAGMENTATION = {"coarse": dict(n_blobs=48, thin=0), "street": dict(n_blobs=400, thin=60), "fine": dict(n_blobs=3000, thin=300)}


def synthetic_labels(index, num_classes, height, width, n_blobs=48, ignore_frac=0.03, thin=0, fragmentation=None):
    if fragmentation is not None:
        n_blobs, thin = FRAGMENTATION[fragmentation]["n_blobs"], FRAGMENTATION[fragmentation]["thin"]
    rng = np.random.RandomState(977 + 31 * int(index))
    ys = rng.randint(0, height, size=n_blobs)
    xs = rng.randint(0, width, size=n_blobs)
    cls = rng.choice(num_classes, size=n_blobs, p=class_prior(num_classes))
    # anisotropic metric -> elongated regions (roads, sky bands)
    ay = rng.uniform(0.5, 2.0, size=n_blobs).astype(np.float32)
    ax = rng.uniform(0.5, 2.0, size=n_blobs).astype(np.float32)
    if n_blobs <= 64:
        yy = np.arange(height, dtype=np.float32)[:, None, None]
        xx = np.arange(width, dtype=np.float32)[None, :, None]
        best = np.full((height, width), np.inf, dtype=np.float32)
        lab = np.zeros((height, width), dtype=np.uint8)
        for b in range(n_blobs):  # streaming arg-min keeps memory at O(H*W)
            d = ay[b] * (yy[:, :, 0] - ys[b]) ** 2 + ax[b] * (xx[:, :, 0] - xs[b]) ** 2
            m = d < best
            best[m] = d[m]
            lab[m] = cls[b]
    else:  # many seeds: nearest seed through a k-d tree on a 1/4-resolution grid (one global anisotropy), blown up 4x
        from scipy.spatial import cKDTree
        sy, sx = float(np.sqrt(ay.mean())), float(np.sqrt(ax.mean()))
        tree = cKDTree(np.stack([ys * sy, xs * sx], axis=1))
        q = 4
        gy, gx = np.meshgrid((np.arange(0, height, q, dtype=np.float32) + 0.5 * q) * sy,
                             (np.arange(0, width, q, dtype=np.float32) + 0.5 * q) * sx, indexing="ij")
        _, nearest = tree.query(np.stack([gy.ravel(), gx.ravel()], axis=1), k=1)
        small = cls[nearest].reshape(gy.shape).astype(np.uint8)
        lab = np.repeat(np.repeat(small, q, axis=0), q, axis=1)[:height, :width].copy()
    for _ in range(thin):  # thin structures: vertical poles / horizontal markings, 2-6 px wide
        c = rng.choice(num_classes, p=class_prior(num_classes))
        wdt = rng.randint(2, 7)
        if rng.rand() < 0.6:
            ln = rng.randint(max(height // 10, 2), max(int(height * 0.6), 3))
            y0, x0 = rng.randint(0, height - ln + 1), rng.randint(0, max(width - wdt, 1))
            lab[y0:y0 + ln, x0:x0 + wdt] = c
        else:
            ln = rng.randint(max(width // 10, 2), max(int(width * 0.5), 3))
            y0, x0 = rng.randint(0, max(height - wdt, 1)), rng.randint(0, width - ln + 1)
            lab[y0:y0 + wdt, x0:x0 + ln] = c
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


def synthetic_batch(indices, num_classes, height, width, label_dtype=torch.uint8, fragmentation=None):
    imgs = torch.stack([synthetic_images(i, height, width) for i in indices])
    labs = torch.stack([synthetic_labels(i, num_classes, height, width, fragmentation=fragmentation) for i in indices]).to(label_dtype)
    return imgs, labs


def label_run_stats(labels, stride=8):
    """(mean run length along x, distinct classes per image) of the labels as a stride-`stride` feature map sees them
    (legacy nearest = every stride-th pixel): what K1's state-free accumulation and slot cache are sensitive to."""
    lab = torch.as_tensor(labels)[..., ::stride, ::stride].numpy()
    changes = (lab[..., 1:] != lab[..., :-1]).sum()
    rows = lab.size // lab.shape[-1]
    runs = changes + rows
    classes = float(np.mean([len(np.unique(l)) for l in lab.reshape(-1, lab.shape[-2], lab.shape[-1])]))
    return float(lab.size / max(runs, 1)), classes
