import numpy as np

from moonshine_b200 import hydra, scenes
from oracle.binding import Oracle


def gpu_render(sc, n_launches=1, **kw):
    img, stats, m = hydra.render_scene(sc, n_launches=n_launches, **kw)
    return img, stats, m


def oracle_render(sc, n_launches=1):
    o = Oracle()
    scenes.load_scene(o, sc)
    img, counts = o.render(sc.lens, *sc.extent, n_launches=n_launches)
    return img, counts, o


def within_tol_fraction(a, b, rel=1e-4):
    """fraction of pixel components with |a-b| <= rel*max(1,|b|) (SURVEY §8c parity definition)."""
    d = np.abs(a[..., :3].astype(np.float64) - b[..., :3].astype(np.float64))
    return float((d <= rel * np.maximum(1.0, np.abs(b[..., :3]))).mean())


def random_rays(seed, n, center, radius, tmax=1e12):
    rs = np.random.RandomState(seed)
    o = (rs.randn(n, 3) * radius + center).astype(np.float32)
    d = rs.randn(n, 3)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = o
    rays[:, 4:7] = d
    rays[:, 7] = tmax
    return rays


def identical_fraction(a, b):
    """fraction of pixels whose four components are the same bits"""
    return float((a.view(np.uint32) == b.view(np.uint32)).all(axis=-1).mean())
