// Host side of the TMA path: CUtensorMap descriptors of the pyramid planes.  The driver entry point
// is fetched through the runtime (no link-time dependency on libcuda).
#include "ctx.h"

namespace sg {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess
            && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int encode_plane_map(sg_ctx *ctx, CUtensorMap *out, const uint8_t *base, int w, int h, int pitch, size_t frame_stride,
                     int frames, int box_w, int box_h, bool swizzle64) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(ctx, SG_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    if (((uintptr_t)base & 15) || (pitch & 15) || (frame_stride & 15) || (box_w & 15) || box_w > 256 || box_h > 256)
        return fail(ctx, SG_ERR_INVALID, "plane not addressable by TMA (base/pitch/stride must be 16-byte multiples, box <= 256)");
    if (swizzle64 && box_w != 64) return fail(ctx, SG_ERR_INVALID, "the 64-byte swizzle needs a 64-byte box row");
    const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)frames};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frame_stride};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, SG_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return SG_OK;
}

// Descriptors whose base is the level-0 plane: the caller's device images or the context's own copy.
int encode_level0_maps(sg_ctx *ctx) {
    const Level &L0 = ctx->lv[0];
    const int frames = ctx->frames_ready > 0 ? ctx->frames_ready : 1;
    if (int r = encode_plane_map(ctx, &ctx->lv[0].map_src, ctx->level0, L0.w, L0.h, ctx->level0_pitch, ctx->level0_stride,
                                 frames, 96, 70)) return r;   // window of the blur-only kernel: 64 + 32 x 64 + 6
    if (int r = encode_plane_map(ctx, &ctx->lv[0].map_fast, ctx->level0, L0.w, L0.h, ctx->level0_pitch, ctx->level0_stride,
                                 frames, 80, 70)) return r;
    int mom_w, mom_h, blur_w, blur_h;
    describe_box_dims(&mom_w, &mom_h, &blur_w, &blur_h);
    if (int r = encode_plane_map(ctx, &ctx->lv[0].map_mom, ctx->level0, L0.w, L0.h, ctx->level0_pitch, ctx->level0_stride,
                                 frames, mom_w, mom_h)) return r;
    if (ctx->p.levels > 1 && ctx->lv[1].fast_resize)
        if (int r = encode_plane_map(ctx, &ctx->lv[1].map_src, ctx->level0, L0.w, L0.h, ctx->level0_pitch, ctx->level0_stride,
                                     frames, ctx->lv[1].tma_src_w, ctx->lv[1].tma_src_h)) return r;
    return SG_OK;
}

}  // namespace sg
