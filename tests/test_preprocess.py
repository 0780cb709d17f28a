"""SURVEY.md 8(f) row 3 — image preprocessing (data/utils/hypervla_interface.py:89-121).

CPU: the oracle's separable anti-aliased resampling machinery is pinned against torch's
`interpolate(antialias=True)` by swapping in the kernels torch implements (Keys cubic, triangle); the lanczos3 kernel
itself and the crop follow TensorFlow's published kernels (TF absent: unpinned, see oracle/preprocess_oracle.py); the
product's host-side span tables equal the oracle's bit for bit.
GPU: the CUDA path (through the C ABI) is bit-exact against the oracle, byte for byte, incl. ragged/edge sizes."""
import numpy as np
import pytest


def _img(seed, h, w):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("shape", [(480, 640), (128, 160), (224, 224), (300, 225)])
@pytest.mark.parametrize("kern", ["bicubic", "bilinear"])
def test_resampling_machinery_matches_torch_antialias(shape, kern):
    torch = pytest.importorskip("torch")
    from oracle import preprocess_oracle as PO
    img = _img(1, *shape)
    k, r = (PO.keys_cubic, 2.0) if kern == "bicubic" else (PO.triangle, 1.0)
    mine = PO.resize_lanczos3(img, 224, k, r)
    ref = torch.nn.functional.interpolate(torch.from_numpy(img.astype(np.float32)).permute(2, 0, 1)[None], size=(224, 224),
                                          mode=kern, antialias=True, align_corners=False)[0].permute(1, 2, 0).numpy()
    assert np.abs(mine - ref).max() < 1e-3          # 0..255 scale, float32 accumulation-order noise


def test_lanczos3_properties():
    from oracle import preprocess_oracle as PO
    # weights sum to one, a 224x224 input is returned unchanged, a constant image stays constant
    for n_in in (224, 256, 480, 640, 100):
        st, w = PO.compute_spans(224, n_in)
        assert np.allclose(w.sum(1), 1.0, atol=1e-5) and st.min() >= 0 and (st + (w != 0).sum(1) <= n_in).all()
    img = _img(2, 224, 224)
    assert np.array_equal(PO.resize_image(img, 224), img)
    const = np.full((480, 640, 3), 77, np.uint8)
    assert np.array_equal(PO.resize_image(const, 224), np.full((224, 224, 3), 77, np.uint8))
    assert np.array_equal(PO.resize_image(const, 224, crop=True), np.full((224, 224, 3), 77, np.uint8))
    # kernel values: 1 at 0, 0 at the integers and beyond the radius
    x = np.array([0.0, 1.0, 2.0, 3.0, 3.5, 0.5], np.float32)
    k = PO.lanczos3(x)
    assert k[0] == 1 and np.abs(k[1:4]).max() < 1e-6 and k[4] == 0 and abs(k[5] - 0.6079271) < 1e-5
    assert PO.round_clip_u8(np.array([0.5, 1.5, 2.5, -3.0, 300.0], np.float32)).tolist() == [0, 2, 2, 0, 255]


def test_host_span_tables_equal_oracle():
    from hvla import preprocess as HP
    from oracle import preprocess_oracle as PO
    for n_in in (224, 256, 480, 512, 640, 37):
        s1, w1 = HP.lanczos3_spans(224, n_in)
        s2, w2 = PO.compute_spans(224, n_in)
        assert np.array_equal(s1, s2) and np.array_equal(w1, w2)


@pytest.mark.gpu
@pytest.mark.parametrize("crop", [False, True])
@pytest.mark.parametrize("shape", [(480, 640), (512, 640), (256, 256), (224, 224), (97, 131)])
def test_gpu_preprocess_bit_exact_vs_oracle(shape, crop):
    import torch
    assert torch.cuda.is_available()
    from hvla.preprocess import BatchedImagePreprocessor
    from oracle import preprocess_oracle as PO
    imgs = np.stack([_img(10 + i, *shape) for i in range(3)])
    pre = BatchedImagePreprocessor(224, crop=crop)
    out = pre(imgs)
    assert out.is_cuda and out.dtype == torch.uint8 and tuple(out.shape) == (3, 224, 224, 3)
    got = out.cpu().numpy()
    for i in range(3):
        assert np.array_equal(got[i], PO.resize_image(imgs[i], 224, crop=crop)), i
    assert pre(imgs[:0]).shape[0] == 0                                    # empty batch
    with pytest.raises(ValueError):
        pre(imgs.astype(np.float32))
    with pytest.raises(ValueError):
        BatchedImagePreprocessor(224, padded_resize=True)


@pytest.mark.gpu
def test_gpu_preprocess_feeds_act():
    """resize on the GPU -> sample_actions on the GPU tensor == sample_actions on the oracle-resized host frames."""
    import torch
    from hvla import config as C, synthetic as S
    from hvla.model import HyperVLA
    from hvla.preprocess import BatchedImagePreprocessor
    from oracle import preprocess_oracle as PO
    m = HyperVLA.from_config(C.default_config(), precision="bf16", params_variant="P1")
    inp = S.make_inputs(4, 2, 2)
    bp, tasks, _ = m.create_tasks(instruction_dict=inp["instruction_dict"], initial_state=inp["initial_state"])
    frames = np.stack([_img(40 + i, 480, 640) for i in range(2)])
    dev = BatchedImagePreprocessor(224)(frames)
    a_dev, _ = m.sample_actions(dev, None, tasks, None, bp)
    host = np.stack([PO.resize_image(f, 224) for f in frames])
    a_host, _ = m.sample_actions(host, None, tasks, None, bp)
    a_dev = a_dev.cpu().numpy() if torch.is_tensor(a_dev) else a_dev
    assert np.array_equal(a_dev, a_host)
