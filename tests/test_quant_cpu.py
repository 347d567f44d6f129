"""CPU: the F16F8 operand format restatement (oracle/quant.py) -- known answers, exact residuals, saturation, block layout, and
the accuracy the format is designed for (value + residual carries ~15 bits; each e4m3 copy 2^-4 relative)."""
import numpy as np
import torch

from oracle import quant


def test_known_answers_and_saturation():
    x = torch.tensor([[0.0, 1.0, -1.0, 0.1, 448.0, 1000.0, -1e5, 2.0 ** -10] + [0.0] * 56])
    h, cross, (hd, v8, r8) = quant.quantize(x, 1.0, 1.0)
    assert h.dtype == torch.float16 and cross.shape == (1, 1, 128) and cross.dtype == torch.uint8
    np.testing.assert_array_equal(hd[0, :8].numpy(), [0.0, 1.0, -1.0, float(np.float16(0.1)), 448.0, 1000.0, -65504.0, 2.0 ** -10])
    np.testing.assert_array_equal(v8[0, :8].numpy(), [0.0, 1.0, -1.0, 0.1015625, 448.0, 448.0, -448.0, 0.0])   # 2^-10 is half of e4m3's smallest subnormal (2^-9): ties to even -> 0
    # e4m3 bytes: 1.0 = 0x38, -1.0 = 0xB8, 448 = 0x7E (largest finite), 0 = 0x00
    assert cross[0, 0, 1].item() == 0x38 and cross[0, 0, 2].item() == 0xB8 and cross[0, 0, 4].item() == 0x7E and cross[0, 0, 0].item() == 0
    # residual of 0.1: (0.1f - fp16(0.1)) * 2^11, rounded to e4m3
    r = (np.float32(0.1) - np.float32(np.float16(0.1))) * np.float32(2048.0)
    assert abs(r8[0, 3].item() * 2048.0 - float(torch.tensor(r).to(torch.float8_e4m3fn).float())) == 0.0


def test_residual_is_exact_and_value_plus_residual_carries_fifteen_bits():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(4, 7, 128, generator=g) * 3.0
    s, t = quant.activation_scales(float(x.abs().max()))
    h, cross, (hd, v8, r8) = quant.quantize(x, s, t)
    # h + (x*s - h) == x*s exactly in float32 (the residual of an fp16 rounding is representable)
    assert torch.equal((h.float() + (x * s - h.float())), x * s)
    full, copy = quant.dequantize(h, cross, s, t)
    rel = ((full - x.double()).abs() / x.double().abs().clamp_min(1e-3)).max().item()
    assert rel < 2.0 ** -14                                    # fp16's 2^-11 residual, known to 2^-4: ~2^-15..2^-16
    rel8 = ((copy - x.double()).abs() / x.double().abs().clamp_min(float(x.abs().max()) * 2.0 ** -8)).max().item()
    assert rel8 <= 2.0 ** -4 + 1e-6                            # one e4m3 rounding


def test_scales_are_powers_of_two_with_the_documented_headroom():
    for amax in (0.37, 1.0, 11.6, 150.0, 3e4):
        s, t = quant.activation_scales(amax)
        assert np.log2(s) == np.floor(np.log2(s)) and np.log2(t) == np.floor(np.log2(t))
        assert amax * s <= 16384.0 * 2 and 448.0 / 8 < amax * t <= 448.0 / 4 + 1e-9      # two binades below saturation
    for wmax in (0.004, 0.3, 2.0):
        s, t = quant.weight_scales(wmax)
        assert 8192.0 < wmax * s <= 16384.0 and 224.0 < wmax * t <= 448.0


def test_block_layout_values_then_residuals_per_64_channels():
    x = torch.arange(128, dtype=torch.float32).reshape(1, 128) / 8.0
    h, cross, _ = quant.quantize(x, 1.0, 1.0)
    assert cross.shape == (1, 2, 128)
    blk0_vals = cross[0, 0, :64].contiguous().view(torch.float8_e4m3fn).float()
    blk1_vals = cross[0, 1, :64].contiguous().view(torch.float8_e4m3fn).float()
    np.testing.assert_allclose(blk0_vals.numpy(), x[0, :64].numpy(), rtol=2.0 ** -4)
    np.testing.assert_allclose(blk1_vals.numpy(), x[0, 64:].numpy(), rtol=2.0 ** -4)
