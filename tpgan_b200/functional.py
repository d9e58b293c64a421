"""Stand-alone (untraced) execution of one conv()/deconv() layer through the C-ABI kernels, used when a layer built by
tpgan_b200.ModificationLayer is called directly on an NCHW CUDA tensor.  Inference only: training goes through the traced
modules of tpgan_b200.D_and_G_model (which fuse bias/residual/activation and implement dgrad/wgrad)."""
from __future__ import annotations

import torch

from . import ops


def conv2d_standalone(mod, x: torch.Tensor, transposed: bool) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError("tpgan_b200 layers run on CUDA tensors only (no CPU fallback)")
    if torch.is_grad_enabled() and (x.requires_grad or mod.weight.requires_grad):
        raise NotImplementedError("stand-alone TCConv2d/TCConvTranspose2d calls are inference-only; wrap the call in "
                                  "torch.no_grad() or use the traced modules of tpgan_b200.D_and_G_model for training")
    k, s, p = mod.kernel_size[0], mod.stride[0], mod.padding[0]
    n, c, h, w = x.shape
    if transposed:
        cout = mod.weight.shape[1]
        op = mod.output_padding[0]
        ho, wo = (h - 1) * s - 2 * p + k + op, (w - 1) * s - 2 * p + k + op
        kind = ops.DECONV_FWD
        if k > 1 and s == 1 and h == 1 and w == 1 and p == 0:   # ConvTranspose on a 1x1 map: a GEMM
            wl = mod.weight.detach().permute(2, 3, 1, 0).reshape(k * k * cout, c, 1, 1).contiguous()
            pw = ops.pack_weights(wl, ops.CONV_FWD)
            out = ops.Act.empty(n, k, k, cout)
            flat = ops.Act(out.buf.view(n, 1, 1, k * k * cout))
            b = None if mod.bias is None else mod.bias.detach().repeat(k * k).contiguous()
            ops.conv2d(ops.CONV_FWD, ops.Act.empty(n, 1, 1, c).from_nchw(x.float()), flat, pw, 1, 1, 0, bias=b)
            return out.to_nchw()
    else:
        cout = mod.weight.shape[0]
        ho, wo = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
        kind = ops.CONV_FWD
    xa = ops.Act.empty(n, h, w, c).from_nchw(x.float())
    out = ops.Act.empty(n, ho, wo, cout)
    pw = ops.pack_weights(mod.weight.detach().contiguous(), kind)
    bias = None
    if mod.bias is not None:
        bias = torch.zeros(ops.round_up(cout, 4), device=x.device)
        bias[:cout] = mod.bias.detach()
    ops.conv2d(kind, xa, out, pw, k, s, p, bias=bias)
    return out.to_nchw()
