"""CPU oracle of the per-step image preprocessing (TEST INFRASTRUCTURE ONLY).

NumPy float32 restatement of `InferenceWrapper._resize_image` (data/utils/hypervla_interface.py:89-121):
    tf.image.resize(image, (S, S), method="lanczos3", antialias=True)            :98-103
    [crop]  tf.image.crop_and_resize(image[None], [[o, o, o+s, o+s]], [0], (S, S)),  s = sqrt(0.9), o = (1-s)/2   :105-119
    tf.cast(tf.clip_by_value(tf.round(image), 0, 255), tf.uint8)                 :120
(`padded_resize` / resize_with_pad, :90-95, is off in every reference config and not covered.)

TensorFlow is a third-party dependency that is absent here (requirements_full_install.txt pins tensorflow==2.15.0),
so its two kernels are restated from their published algorithm:
  * ScaleAndTranslate (core/kernels/image/scale_and_translate_op.cc): per output index x, sample_f = (x+0.5)/scale;
    kernel_scale = max(1/scale, 1) (antialias); span = [ceil(sample_f - 3*ks - 0.5), floor(sample_f + 3*ks - 0.5)]
    clamped to the image; weight = lanczos3(|source + 0.5 - sample_f| / ks), normalised by the span's sum; rows
    (vertical) are gathered first into a float32 intermediate, then columns; taps accumulate in span order.
  * CropAndResize (core/kernels/image/crop_and_resize_op.cc), bilinear, extrapolation_value 0.
The reference has no fixtures for this step and TF cannot run here: PARITY UNPINNED for the TF internals.  What pins
the span/weight machinery: tests/test_preprocess.py swaps in the Keys cubic / triangle kernels and compares with
torch.nn.functional.interpolate(antialias=True), which implements the same separable anti-aliased resampling.
"""
import numpy as np

F = np.float32


def lanczos3(x):
    """LanczosKernelFunc(radius=3): 0 beyond the radius, 1 at |x| <= 1e-3, else 3 sin(pi x) sin(pi x / 3) / (pi^2 x^2)."""
    x = np.abs(x.astype(F))
    pi = F(3.14159265359)
    safe = np.where(x <= F(1e-3), F(1), x)
    v = F(3) * np.sin(pi * safe).astype(F) * np.sin(pi * safe / F(3)).astype(F) / (pi * pi * safe * safe)
    v = np.where(x <= F(1e-3), F(1), v)
    return np.where(x > F(3), F(0), v).astype(F)


def keys_cubic(x):
    """Keys cubic (a = -0.5), radius 2: the kernel of TF's 'bicubic' and of torch's antialiased bicubic."""
    x = np.abs(x.astype(F))
    a = ((F(1.5) * x - F(2.5)) * x) * x + F(1)
    b = ((F(-0.5) * x + F(2.5)) * x - F(4)) * x + F(2)
    return np.where(x >= F(2), F(0), np.where(x >= F(1), b, a)).astype(F)


def triangle(x):
    x = np.abs(x.astype(F))
    return np.maximum(F(0), F(1) - x).astype(F)


def compute_spans(out_size, in_size, kernel=lanczos3, radius=3.0, antialias=True):
    """ComputeSpansCore: (starts int32 [out], weights float32 [out, span_size])."""
    scale = F(out_size) / F(in_size)
    inv_scale = F(1) / scale
    ks = max(inv_scale, F(1)) if antialias else F(1)
    span_size = min(2 * int(np.ceil(F(radius) * ks)) + 1, in_size)
    starts = np.zeros(out_size, np.int32)
    weights = np.zeros((out_size, span_size), F)
    one_over = F(1) / ks
    for x in range(out_size):
        sample = F(F(x) + F(0.5)) * inv_scale
        if sample < 0 or sample > in_size:
            continue
        s = int(np.ceil(sample - F(radius) * ks - F(0.5)))
        e = int(np.floor(sample + F(radius) * ks - F(0.5)))
        s = min(max(s, 0), in_size - 1)
        e = min(max(e, 0), in_size - 1) + 1
        src = np.arange(s, e).astype(F)
        w = kernel(np.abs((src + F(0.5) - sample) * one_over))
        tot = F(0)
        for v in w:                       # float32 running sum in span order, as the C++ loop does
            tot = F(tot + v)
        if abs(tot) >= F(1000) * np.finfo(F).tiny:
            w = (w * (F(1) / tot)).astype(F)
            weights[x, :len(w)] = w
        starts[x] = s
    return starts, weights


def gather(img, starts, weights, axis):
    """GatherRows / GatherColumns: out[i] = sum_k weights[i, k] * img[starts[i] + k] (float32, taps in order;
    taps past the image edge have zero weight and are skipped like the C++ real_span_size clamp)."""
    img = np.moveaxis(img.astype(F), axis, 0)
    n_in = img.shape[0]
    out = np.zeros((len(starts),) + img.shape[1:], F)
    for k in range(weights.shape[1]):
        idx = np.minimum(starts + k, n_in - 1)
        valid = (starts + k) < n_in
        w = np.where(valid, weights[:, k], F(0)).reshape((-1,) + (1,) * (img.ndim - 1))
        out = (out + (w * img[idx]).astype(F)).astype(F)
    return np.moveaxis(out, 0, axis)


def resize_lanczos3(image_u8, size, kernel=lanczos3, radius=3.0):
    """tf.image.resize(..., method='lanczos3', antialias=True) of one HxWxC image -> float32 (size,size,C)."""
    H, W = image_u8.shape[:2]
    sy, wy = compute_spans(size, H, kernel, radius)
    sx, wx = compute_spans(size, W, kernel, radius)
    tmp = gather(image_u8.astype(F), sy, wy, 0)          # rows first (intermediate is [out_h, in_w])
    return gather(tmp, sx, wx, 1)


def crop_and_resize_center(img_f32, size, scale=np.sqrt(0.9)):
    """tf.image.crop_and_resize with the single box [o, o, o+s, o+s] (hypervla_interface.py:105-119), bilinear."""
    H, W = img_f32.shape[:2]
    o = (1 - scale) / 2
    y1, x1, y2, x2 = F(o), F(o), F(o + scale), F(o + scale)     # the box tensor is float32
    hs = (y2 - y1) * F(H - 1) / F(size - 1)
    ws = (x2 - x1) * F(W - 1) / F(size - 1)
    out = np.zeros((size, size, img_f32.shape[2]), F)
    in_y = (y1 * F(H - 1) + np.arange(size).astype(F) * hs).astype(F)
    in_x = (x1 * F(W - 1) + np.arange(size).astype(F) * ws).astype(F)
    ty, by = np.floor(in_y).astype(int), np.ceil(in_y).astype(int)
    lx, rx = np.floor(in_x).astype(int), np.ceil(in_x).astype(int)
    yl = (in_y - ty.astype(F)).astype(F)[:, None, None]
    xl = (in_x - lx.astype(F)).astype(F)[None, :, None]
    ok_y, ok_x = (in_y >= 0) & (in_y <= H - 1), (in_x >= 0) & (in_x <= W - 1)
    ty, by, lx, rx = [np.clip(v, 0, n - 1) for v, n in ((ty, H), (by, H), (lx, W), (rx, W))]
    tl, tr = img_f32[ty][:, lx], img_f32[ty][:, rx]
    bl, br = img_f32[by][:, lx], img_f32[by][:, rx]
    top = (tl + ((tr - tl).astype(F) * xl).astype(F)).astype(F)
    bot = (bl + ((br - bl).astype(F) * xl).astype(F)).astype(F)
    out = (top + ((bot - top).astype(F) * yl).astype(F)).astype(F)
    out[~ok_y, :, :] = 0
    out[:, ~ok_x, :] = 0
    return out


def round_clip_u8(img_f32):
    """tf.cast(tf.clip_by_value(tf.round(x), 0, 255), tf.uint8); tf.round is round-half-to-even (= np.rint)."""
    return np.clip(np.rint(img_f32), 0, 255).astype(np.uint8)


def resize_image(image_u8, image_size=224, crop=False):
    """InferenceWrapper._resize_image for one image (H,W,3) uint8 -> (image_size,image_size,3) uint8."""
    img = resize_lanczos3(image_u8, image_size)
    if crop:
        img = crop_and_resize_center(img, image_size)
    return round_clip_u8(img)
