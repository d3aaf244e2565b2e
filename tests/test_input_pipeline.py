"""Input pipeline: the oracle restatement (oracle/image_oracle.py) against the golden vectors the
reference's PreprocessRGB + Pillow produced (CPU), and the device kernels against the oracle — bit-exact for the uint8
image AND for the float tensor (GPU)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import image_oracle

GOLDEN = Path(__file__).resolve().parent / "golden"


def _cases(fx):
    for key in fx.files:
        if key.endswith("_img"):
            base = key[:-4]
            size, mode, _ = base.split("_")
            yield base, int(size[1:]), mode, fx[key], fx[base + "_ref"], fx[base + "_u8"]


def test_oracle_matches_reference_golden():
    fx = np.load(GOLDEN / "image_preprocess.npz")
    n = 0
    for base, size, mode, img, ref, u8 in _cases(fx):
        got, got_u8 = image_oracle.preprocess_rgb(img, size, fx["image_mean"], fx["image_std"], 1 / 255, mode)
        assert np.array_equal(got_u8, u8), base            # == Pillow's Image.resize(BICUBIC), byte for byte
        assert np.array_equal(got, ref), base              # == the reference's PreprocessRGB output tensor
        n += 1
    assert n >= 10
    stats = {k: fx["stat_" + k] for k in ("min", "max", "mean", "std")}
    assert np.array_equal(image_oracle.action_normalize(fx["action_in"], stats, True), fx["action_quantile"])
    assert np.array_equal(image_oracle.action_normalize(fx["action_in"], stats, False), fx["action_meanstd"])


def test_oracle_resize_equals_pillow_when_available():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(3)
    for (H, W, size) in ((97, 131, 48), (64, 64, 96), (50, 200, 30)):
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        sq = image_oracle.expand2square(img, (10, 20, 30))
        ref = np.asarray(Image.fromarray(sq).resize((size, size), resample=Image.BICUBIC))
        assert np.array_equal(image_oracle.pil_bicubic_resize(sq, size), ref)


def test_host_coefficient_tables_match_oracle():
    """The product's host-side table builder (dexbotic_b200/input_pipeline.py) == the oracle's restatement of Pillow's
    precompute_coeffs / normalize_coeffs_8bpc (no GPU needed)."""
    from dexbotic_b200.input_pipeline import resample_coefficients
    for in_size, out_size in ((640, 224), (480, 224), (100, 64), (224, 384), (37, 64)):
        kk, bounds, ksize = resample_coefficients(in_size, out_size)
        okk, obounds = image_oracle.precompute_coeffs(in_size, out_size)
        assert ksize == okk.shape[1] and np.array_equal(kk, okk) and np.array_equal(bounds, obounds)


@pytest.mark.gpu
def test_image_preprocess_kernels_bit_exact():
    from dexbotic_b200.input_pipeline import ImagePreprocessor
    fx = np.load(GOLDEN / "image_preprocess.npz")
    for base, size, mode, img, ref, u8 in _cases(fx):
        pp = ImagePreprocessor(size=size, image_mean=fx["image_mean"], image_std=fx["image_std"], image_pad_mode=mode)
        frames = torch.from_numpy(np.stack([img, img[::-1].copy()])).cuda()           # batch of 2: the frame and its flip
        out, got_u8 = pp(frames, return_uint8=True)
        assert np.array_equal(got_u8[0].cpu().numpy(), u8), base
        assert np.array_equal(out[0].cpu().numpy(), ref), base
        flip_ref, flip_u8 = image_oracle.preprocess_rgb(img[::-1].copy(), size, fx["image_mean"], fx["image_std"], 1 / 255, mode)
        assert np.array_equal(got_u8[1].cpu().numpy(), flip_u8) and np.array_equal(out[1].cpu().numpy(), flip_ref)
    # production geometry: a batch of 32 VGA frames -> 224 x 224, bf16 output = one rounding of the fp32 tensor
    rng = np.random.default_rng(11)
    frames = rng.integers(0, 256, (32, 480, 640, 3), dtype=np.uint8)
    pp = ImagePreprocessor(size=224)
    t = torch.from_numpy(frames).cuda()
    out32 = pp(t)
    out16 = pp(t, dtype=torch.bfloat16)
    for b in (0, 13, 31):
        ref, _ = image_oracle.preprocess_rgb(frames[b], 224, pp.mean, pp.std, 1 / 255, "mean")
        assert np.array_equal(out32[b].cpu().numpy(), ref)
    assert torch.equal(out16, out32.to(torch.bfloat16))


@pytest.mark.gpu
def test_action_normalize_kernel_bit_exact():
    from dexbotic_b200.input_pipeline import ActionNormalizer
    fx = np.load(GOLDEN / "image_preprocess.npz")
    stats = {k: fx["stat_" + k] for k in ("min", "max", "mean", "std")}
    a = torch.from_numpy(fx["action_in"]).cuda()
    for q, key in ((True, "action_quantile"), (False, "action_meanstd")):
        got = ActionNormalizer(stats, use_quantiles=q)(a).cpu().numpy()
        assert got.dtype == np.float32 and np.array_equal(got, fx[key])
