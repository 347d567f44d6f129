"""Oracle (test infrastructure; never imported by the product): CPU restatement of the F16F8 operand format of the 2-unit
convolution arithmetic (text-detection-ctpn_b200/csrc/common.cuh).  This is this repo's own format -- the
reference computes in float32 (lib/networks/network.py:160-183) -- so there is nothing in the reference to pin it against; it is
pinned against torch's float16 / float8_e4m3fn conversions on the CPU (tests/test_quant_cpu.py) and used by the GPU checks to
build operands and to decode outputs (tests/gpu_checks.py).

A float32 value a is carried as
    h  = fp16_rn(a * s)                    (plane 0; saturating at +-65504)
    v8 = e4m3_rn(a * t)                    (plane 1, first 64 bytes of the pixel's 128-byte block of a 64-channel group)
    r8 = e4m3_rn((a * s - h) * 2^11 t / s) (plane 1, second 64 bytes; the residual is exact in float32)
with per-tensor powers of two s, t; e4m3 saturates at +-448.  Weight rows store the residual half FIRST."""
import numpy as np
import torch

RESIDUAL_GAIN = 2048.0


def pow2_floor(v):
    return float(2.0 ** np.floor(np.log2(v)))


def quantize(x, s, t):
    """float32 tensor [..., C] (C % 64 == 0) -> (fp16 plane, uint8 cross plane [..., C/64, 128], dequantised (h, v8, r8) float64)."""
    x = x.to(torch.float32)
    h = (x * s).clamp(-65504.0, 65504.0).to(torch.float16)
    r = x * s - h.float()
    v8 = (x * t).clamp(-448, 448).to(torch.float8_e4m3fn)
    r8 = (r * (RESIDUAL_GAIN * t / s)).clamp(-448, 448).to(torch.float8_e4m3fn)
    lead, C = x.shape[:-1], x.shape[-1]
    cross = torch.cat([v8.view(torch.uint8).reshape(lead + (C // 64, 64)), r8.view(torch.uint8).reshape(lead + (C // 64, 64))], dim=-1)
    return h, cross.contiguous(), (h.double() / s, v8.double() / t, r8.double() / (RESIDUAL_GAIN * t))


def dequantize(h, cross, s, t):
    """(fp16 plane [..., C], uint8 cross plane [..., C/64, 128]) -> (value + residual, e4m3 value copy), float64 [..., C]."""
    lead, C = h.shape[:-1], h.shape[-1]
    v8 = cross[..., :64].contiguous().view(torch.float8_e4m3fn).double().reshape(lead + (C,)) / t
    r8 = cross[..., 64:].contiguous().view(torch.float8_e4m3fn).double().reshape(lead + (C,)) / (RESIDUAL_GAIN * t)
    return h.double() / s + r8, v8


def activation_scales(amax):
    """What ctpn_net's calibration derives from a layer's maximum |activation| (net.cu::calibrate_f16f8)."""
    s = pow2_floor(16384.0 / amax) if amax > 16384.0 else 1.0
    return s, pow2_floor(448.0 / amax) / 4.0


def weight_scales(wmax):
    """net.cu::upload_packed_f16f8."""
    return pow2_floor(16384.0 / wmax), pow2_floor(448.0 / wmax)
