"""Deterministic synthetic weights / inputs (numpy PCG64: identical on every machine, unlike torch's default init
which depends on call order).  Used by bench.py, smoke() and the tests so the CUDA path, the oracle and the
golden fixtures all see the same numbers without shipping 200 MB of weights (SURVEY.md 8d "Synthetic inputs").
There is no network in this environment, so pretrained checkpoints are never available."""
import math
import zlib

import numpy as np
import torch


def _rng(seed, name):
    return np.random.default_rng([seed, zlib.crc32(name.encode())])


def synth_tensor(name, shape, seed):
    """One parameter/buffer, distribution chosen from its name and rank."""
    g = _rng(seed, name)
    shape = tuple(shape)
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.int64)
    if leaf == "running_var":
        return torch.from_numpy(g.uniform(0.5, 1.5, shape).astype(np.float32))
    if leaf == "running_mean":
        return torch.from_numpy((0.1 * g.standard_normal(shape)).astype(np.float32))
    if "pos_embed" in name:
        return torch.from_numpy((0.02 * g.standard_normal(shape)).astype(np.float32))
    if len(shape) >= 2:  # Linear [out,in] / Conv [out,in,kh,kw]: variance-preserving
        fan_in = int(np.prod(shape[1:]))
        return torch.from_numpy((g.standard_normal(shape) / math.sqrt(fan_in)).astype(np.float32))
    if leaf == "weight":  # norm scales
        return torch.from_numpy((1.0 + 0.1 * g.standard_normal(shape)).astype(np.float32))
    return torch.from_numpy((0.1 * g.standard_normal(shape)).astype(np.float32))  # biases


def synth_state_dict(template, seed=1234):
    """template: mapping name -> tensor (only shapes are used).  Returns name -> CPU tensor."""
    return {k: synth_tensor(k, v.shape, seed) for k, v in template.items()}


def synth_images(n, h=480, w=640, seed=20240000):
    """Gray images in [0,1], [n,1,h,w]: ITU-601 gray of uniform uint8 RGB noise, /255."""
    g = np.random.default_rng(seed)
    rgb = g.integers(0, 256, size=(n, 3, h, w), dtype=np.uint8).astype(np.float32)
    gray = 0.299 * rgb[:, 0] + 0.587 * rgb[:, 1] + 0.114 * rgb[:, 2]
    # low-pass so the two images of a pair are correlated shifted copies (gives structured matches)
    return torch.from_numpy((gray / 255.0)[:, None].astype(np.float32))


def synth_pair_images(n, h=480, w=640, seed=20240000, shift=(8, 16)):
    """image1 = image0 rolled by `shift` pixels + noise: random-init features then produce peaked, well-separated
    dual-softmax maxima, like a real overlapping pair."""
    g = np.random.default_rng(seed)
    base = g.random((n, 1, h, w), dtype=np.float32)
    k = 8
    base = base.reshape(n, 1, h // k, k, w // k, k).mean(axis=(3, 5))      # blocky texture
    base = np.repeat(np.repeat(base, k, axis=2), k, axis=3)
    img0 = base + 0.05 * g.standard_normal(base.shape).astype(np.float32)
    img1 = np.roll(base, shift, axis=(2, 3)) + 0.05 * g.standard_normal(base.shape).astype(np.float32)
    img0 = np.clip((img0 - img0.min()) / (img0.max() - img0.min()), 0, 1)
    img1 = np.clip((img1 - img1.min()) / (img1.max() - img1.min()), 0, 1)
    return torch.from_numpy(img0.astype(np.float32)), torch.from_numpy(img1.astype(np.float32))


def synth_mapfree_images(n, seed=20240000):
    """Map-free inputs: matcher images [n,1,720,544] gray and regression images [n,3,360,270] (bilinear
    resize of the same texture, per-channel gain, roughly zero-mean like the ImageNet-normalised reference input)."""
    i0, i1 = synth_pair_images(n, h=720, w=544, seed=seed)
    gain = torch.tensor([1.0, 0.9, 1.1]).view(1, 3, 1, 1)
    r0 = torch.nn.functional.interpolate(i0, size=(360, 270), mode="bilinear", align_corners=False).repeat(1, 3, 1, 1)
    r1 = torch.nn.functional.interpolate(i1, size=(360, 270), mode="bilinear", align_corners=False).repeat(1, 3, 1, 1)
    return i0, i1, (r0 * gain - 0.45).contiguous(), (r1 * gain - 0.45).contiguous()


def mapfree_intrinsics(n=1):
    """K of the map-free matcher images (f 590, principal point at the image centre of 544 x 720)."""
    K = torch.tensor([[590.0, 0, 272.0], [0, 590.0, 360.0], [0, 0, 1.0]])
    return K[None].repeat(n, 1, 1)


def mp3d_intrinsics(n=1):
    """K of the Matterport pairs: f 517.97, c (320, 240)  (mp3d_loftr/src/utils/dataset.py:201-211)."""
    K = torch.tensor([[517.97, 0, 320.0], [0, 517.97, 240.0], [0, 0, 1.0]])
    return K[None].repeat(n, 1, 1)


def two_view_geometry(P, N, seed=5, noise=1e-3, outlier_frac=0.2):
    """Config-5 solver inputs: X ~ U([-1,1]^2 x [3,5]), rotation <= 30 deg, unit t, normalised
    coordinates, gaussian noise, uniform outliers, w ~ U(0,1).  Returns pts1, pts2 [P,N,2], w [P,N], R [P,3,3], t [P,3]."""
    g = np.random.default_rng(seed)
    X = np.concatenate([g.uniform(-1, 1, (P, N, 2)), g.uniform(3, 5, (P, N, 1))], -1)
    ax = g.standard_normal((P, 3))
    ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    ang = g.uniform(0.05, np.pi / 6, (P, 1))
    Kx = np.zeros((P, 3, 3))
    Kx[:, 0, 1], Kx[:, 0, 2], Kx[:, 1, 0] = -ax[:, 2], ax[:, 1], ax[:, 2]
    Kx[:, 1, 2], Kx[:, 2, 0], Kx[:, 2, 1] = -ax[:, 0], -ax[:, 1], ax[:, 0]
    s, c = np.sin(ang)[:, :, None], np.cos(ang)[:, :, None]
    R = np.eye(3)[None] + s * Kx + (1 - c) * (Kx @ Kx)
    t = g.standard_normal((P, 3))
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    X2 = X @ R.transpose(0, 2, 1) + t[:, None, :]
    p1 = X[..., :2] / X[..., 2:]
    p2 = X2[..., :2] / X2[..., 2:]
    p1 = p1 + noise * g.standard_normal(p1.shape)
    p2 = p2 + noise * g.standard_normal(p2.shape)
    nout = int(outlier_frac * N)
    if nout:
        p2[:, :nout] = g.uniform(-0.4, 0.4, (P, nout, 2))
    w = g.uniform(0, 1, (P, N))
    f = lambda a: torch.from_numpy(a.astype(np.float32))  # noqa: E731
    return f(p1), f(p2), f(w), f(R), f(t)
