"""Does cuTensorMapEncodeTiled accept a box wider than the tensor's innermost extent (OOB columns zero-filled)?"""
import torch
from cuda.bindings import driver as cu

torch.zeros(1, device="cuda")
x = torch.zeros(2, 300, 3 * 12 * 40, device="cuda", dtype=torch.float16)
H, HD, T, B = 12, 40, 300, 2
dims = [cu.cuuint64_t(HD), cu.cuuint64_t(3 * H), cu.cuuint64_t(T), cu.cuuint64_t(B)]
strides = [cu.cuuint64_t(HD * 2), cu.cuuint64_t(3 * H * HD * 2), cu.cuuint64_t(3 * H * HD * T * 2)]
for box0 in (40, 48, 64):
    box = [cu.cuuint32_t(box0), cu.cuuint32_t(1), cu.cuuint32_t(128), cu.cuuint32_t(1)]
    es = [cu.cuuint32_t(1)] * 4
    r = cu.cuTensorMapEncodeTiled(cu.CUtensorMapDataType.CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, x.data_ptr(), dims, strides, box, es,
                                  cu.CUtensorMapInterleave.CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  cu.CUtensorMapSwizzle.CU_TENSOR_MAP_SWIZZLE_128B,
                                  cu.CUtensorMapL2promotion.CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  cu.CUtensorMapFloatOOBfill.CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
    print("box0", box0, "->", r[0])
