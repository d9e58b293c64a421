"""Probe (GPU box): does cuTensorMapEncodeTiled accept a map whose dim-1 stride (16 B) is smaller than the dim-0 extent
(32 floats = 128 B), i.e. overlapping 8-pixel windows of a 4-channel NHWC image?  Prints the CUresult."""
import torch
from cuda import cuda

torch.zeros(1, device="cuda")
x = torch.zeros((2, 134, 136, 4), device="cuda")
W, H, N = 128, 128, 2
dims = [cuda.cuuint64_t(32), cuda.cuuint64_t(W), cuda.cuuint64_t(H + 6), cuda.cuuint64_t(N)]
strides = [cuda.cuuint64_t(16), cuda.cuuint64_t(136 * 16), cuda.cuuint64_t(134 * 136 * 16)]
box = [cuda.cuuint32_t(32), cuda.cuuint32_t(64), cuda.cuuint32_t(1), cuda.cuuint32_t(1)]
estr = [cuda.cuuint32_t(1)] * 4
for swz in (cuda.CUtensorMapSwizzle.CU_TENSOR_MAP_SWIZZLE_128B, cuda.CUtensorMapSwizzle.CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
            if hasattr(cuda.CUtensorMapSwizzle, "CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B") else cuda.CUtensorMapSwizzle.CU_TENSOR_MAP_SWIZZLE_128B):
    r = cuda.cuTensorMapEncodeTiled(cuda.CUtensorMapDataType.CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x.data_ptr(), dims, strides, box, estr,
                                    cuda.CUtensorMapInterleave.CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                    cuda.CUtensorMapL2promotion.CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                    cuda.CUtensorMapFloatOOBfill.CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
    print(swz, r[0])
