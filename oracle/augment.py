"""TEST INFRASTRUCTURE (oracle): numpy restatement of the reference's image augmentation closure
(core/carla_agent.py:527-579) for explicit random parameters -- the checker of `cdra_augment`, never a product path.

Each op cites the reference lines it follows.  [lib] marks TF semantics restated from the published op definitions
(tf.image.adjust_brightness / adjust_contrast / adjust_saturation / adjust_hue, depthwise_conv2d SAME,
tf.image.resize NEAREST with half-pixel centres).  Parity unpinned against TensorFlow itself (TF is not installable
here); pinned by closed-form cases in tests/test_augment.py.
"""
import numpy as np

F = np.float32


def hash32(seed, a, b):
    """Counter-based per-pixel draw shared with csrc/augment.cuh (`lowbias32` finaliser over a mixed key)."""
    with np.errstate(over='ignore'):
        x = (np.uint32(seed) ^ (np.asarray(a, np.uint32) * np.uint32(0x9E3779B1)) ^ (np.asarray(b, np.uint32) * np.uint32(0x85EBCA77))).astype(np.uint32)
        x ^= x >> np.uint32(16); x = (x * np.uint32(0x7feb352d)).astype(np.uint32)
        x ^= x >> np.uint32(15); x = (x * np.uint32(0x846ca68b)).astype(np.uint32)
        x ^= x >> np.uint32(16)
    return x


def rgb_to_hsv(c):
    mx, mn = c.max(-1), c.min(-1)
    d = mx - mn
    s = np.where(mx > 0, d / np.where(mx > 0, mx, 1), 0).astype(F)
    dd = np.where(d > 0, d, 1)
    r, g, b = c[..., 0], c[..., 1], c[..., 2]
    h = np.where(mx == r, np.mod((g - b) / dd, 6), np.where(mx == g, (b - r) / dd + 2, (r - g) / dd + 4))
    h = np.where(d > 0, h, 0) / 6
    return h.astype(F), s, mx.astype(F)


def hsv_to_rgb(h, s, v):          # [lib] tf.image.hsv_to_rgb
    dh = h * F(6)
    dr = np.clip(np.abs(dh - 3) - 1, 0, 1); dg = np.clip(2 - np.abs(dh - 2), 0, 1); db = np.clip(2 - np.abs(dh - 4), 0, 1)
    return np.stack([((d - 1) * s + 1) * v for d in (dr, dg, db)], -1).astype(F)


def color_jitter(x, brightness, contrast, saturation, hue):
    """rl/augmentations/simclr.py:44-50 with the four draws given: tf_brightness (augmentations.py:111-112) adds delta;
    tf_contrast (:107-108) scales around the per-frame channel mean [lib]; tf_saturation (:103-104) scales S in HSV and
    clips it [lib]; tf_hue (:115-116) rotates H [lib]; then clip to [0, 1] (`original=True`)."""
    mean = x.mean(axis=(-3, -2), keepdims=True, dtype=np.float64).astype(F) + F(brightness)
    x = (x + F(brightness) - mean) * F(contrast) + mean
    h, s, v = rgb_to_hsv(x)
    x = hsv_to_rgb(h, np.clip(s * F(saturation), 0, 1).astype(F), v)
    h, s, v = rgb_to_hsv(x)
    h = h + F(hue); h = h - np.floor(h)
    return np.clip(hsv_to_rgb(h.astype(F), s, v), 0, 1).astype(F)


def blur(x, kernel):
    """tf_gaussian_blur (augmentations.py:193-207): depthwise SAME cross-correlation with the given [size, size, 3] kernel."""
    size = kernel.shape[0]; r = size // 2
    H, W = x.shape[-3], x.shape[-2]
    xp = np.zeros(x.shape[:-3] + (H + 2 * r, W + 2 * r, 3), F); xp[..., r:r + H, r:r + W, :] = x
    out = np.zeros_like(x)
    for i in range(size):
        for j in range(size):
            out += kernel[i, j].astype(F) * xp[..., i:i + H, j:j + W, :]
    return out


def _pixel_ids(frames, H, W):
    f = np.arange(frames, dtype=np.uint32)[:, None]
    b = (np.arange(H * W, dtype=np.uint32) * np.uint32(16))[None, :]
    return f, b


def salt_pepper(x, seed, amount):
    """tf_salt_and_pepper_batch (augmentations.py:176-191): select with p = amount / 10, salt with p = 1/2, all 3 channels."""
    frames, H, W = x.shape[0], x.shape[1], x.shape[2]
    f, b = _pixel_ids(frames, H, W)
    sel = (hash32(seed, f, b) >> np.uint32(8)) < np.uint32(F(amount) * F(0.1) * F(16777216.0))
    nz = ((hash32(seed, f, b + np.uint32(1)) >> np.uint32(8)) < np.uint32(8388608)).astype(F)
    sel = sel.reshape(frames, H, W, 1); nz = nz.reshape(frames, H, W, 1)
    return np.where(sel, nz, x).astype(F)


def gaussian_noise(x, seed, amount, std):
    """tf_gaussian_noise_batch (augmentations.py:147-157): x + clip(select * N(0, std), 0, 1), select per pixel."""
    frames, H, W = x.shape[0], x.shape[1], x.shape[2]
    f, b = _pixel_ids(frames, H, W)
    sel = ((hash32(seed, f, b + np.uint32(2)) >> np.uint32(8)) < np.uint32(F(amount) * F(16777216.0))).reshape(frames, H, W, 1)
    n = np.zeros((frames, H * W, 3), F)
    for k in range(3):
        u1 = ((hash32(seed, f, b + np.uint32(3 + k)) >> np.uint32(8)).astype(F) + F(1)) * F(1.0 / 16777216.0)
        u2 = (hash32(seed, f, b + np.uint32(6 + k)) >> np.uint32(8)).astype(F) * F(1.0 / 16777216.0)
        n[..., k] = np.sqrt(F(-2) * np.log(u1)) * np.cos(F(6.28318530717958647692) * u2) * F(std)
    return (x + np.where(sel, np.clip(n.reshape(frames, H, W, 3), 0, 1), 0)).astype(F)


def normalize(x, group, eps):
    """tf_normalize_batch (augmentations.py:253-263): per sample (= `group` consecutive frames) x -= min; x /= max + eps."""
    frames = x.shape[0]
    out = x.copy()
    for g in range(0, frames, group):
        blk = out[g:g + group]
        mn = blk.min(); mx = blk.max()
        out[g:g + group] = (blk - mn) / ((mx - mn) + F(eps))
    return out.astype(F)


def grid_mask(H, W, size, cells):
    """tf.image.resize(..., NEAREST) [lib, half-pixel centres]: pixel (y, x) takes grid cell floor((i + 0.5) * size / extent)."""
    cy = np.minimum(((np.arange(H, dtype=F) + F(0.5)) * F(size) / F(H)).astype(np.int32), size - 1)
    cx = np.minimum(((np.arange(W, dtype=F) + F(0.5)) * F(size) / F(W)).astype(np.int32), size - 1)
    return cells.reshape(size, size)[cy[:, None], cx[None, :]]


def augment(images, p, dropout_mask=None):
    """The whole closure for explicit parameters `p` (an object with the fields of cdra_augment_params).
    images: [frames, H, W, 3] uint8 or float32.  Order: core/carla_agent.py:546-574."""
    x = images.astype(F) / F(255) if images.dtype == np.uint8 else images.astype(F)
    frames, H, W, _ = x.shape
    if p.jitter:
        x = color_jitter(x, p.brightness, p.contrast, p.saturation, p.hue)
    if p.blur_size:
        x = blur(x, np.asarray(list(p.blur_kernel)[:p.blur_size * p.blur_size * 3], F).reshape(p.blur_size, p.blur_size, 3))
    if p.salt_pepper:
        x = salt_pepper(x, p.seed, p.sp_amount)
    if p.gauss_noise:
        x = gaussian_noise(x, p.seed, p.gn_amount, p.gn_std)
    if p.normalize:
        x = normalize(x, p.group, p.eps)
    if p.cutout_size:        # the FIRST image's mask multiplies the whole batch (`[0]` after the resize, augmentations.py:66-68)
        cells = np.ones(p.cutout_size ** 2, F); cells[p.cutout_cell] = 0
        x = x * grid_mask(H, W, p.cutout_size, cells)[None, :, :, None]
    if p.dropout_size:       # same quirk (augmentations.py:90-92)
        x = x * grid_mask(H, W, p.dropout_size, np.asarray(dropout_mask, F).ravel())[None, :, :, None]
    return x.astype(F)
