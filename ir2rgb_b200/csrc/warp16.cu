// warp16.cu -- the bilinear flow warp on 16-bit storage (fp16 / bf16), forward, for sm_100a.
//
// Replaces, each in ONE pass over 16-bit tensors,
//   mode RESAMPLE2D : fp16_resample2d (reference models/flownet2_pytorch/models.py:22-28) =
//                     `Resample2d()(img.float(), flow.float()).half()` -- two widening copies, the fp32 kernel and a
//                     narrowing copy in the reference (4 kernels, 44 B per pixel and channel-set instead of 20);
//   mode GRIDSAMPLE : Model.resample with opt['fp16'] on 16-bit tensors (reference models/base_model.py:123-136):
//                     get_grid in the flow's dtype, two scalar divides, cat and add evaluated in that dtype, then
//                     `F.grid_sample(image.float(), grid.float(), 'bilinear', 'border').half()`.
// The arithmetic between the loads and the store is the fp32 warp's (warp_rows.cuh: pix_prep / pix_blend, shared, not
// copied), so RESAMPLE2D is bit-identical to the cast chain around flowops_warp_fwd, and GRIDSAMPLE reproduces the
// roundings of the 16-bit grid arithmetic where the reference performs them.
//
// Same structure as warp_rows_kernel: a thread walks `rows` consecutive rows of one column with the next row's flow in
// flight, block-uniform batch index, 32-bit corner offsets.  Algorithmic bytes per pixel: 2*(2C+2).
#include "warp_rows.cuh"

namespace flowops {

struct Warp16Args {
    WarpArgs g;                 // geometry and GRIDSAMPLE constants (its fp32 tensor pointers are unused)
    const unsigned short *img;  // [B,C,H,W]
    const unsigned short *flow; // [B,2,H,W], pixels
    unsigned short *out;        // [B,C,H,W]
};

template <typename T, int CT, int MODE>
__device__ __forceinline__ void pix_gather16(float (&v)[CT][4], const PixPrep<MODE> &q, const unsigned short *__restrict__ src,
                                             unsigned hw)
{
    const unsigned short *pt = src + q.o_t, *pb = src + q.o_b;
#pragma unroll
    for (int c = 0; c < CT; ++c) {
        if constexpr (MODE == FLOWOPS_WARP_RESAMPLE2D) {
            v[c][0] = Io16<T>::to_float(__ldg(pt)); v[c][2] = Io16<T>::to_float(__ldg(pb));
            v[c][1] = q.ex ? Io16<T>::to_float(__ldg(pt + 1)) : 0.f;
            v[c][3] = q.ex ? Io16<T>::to_float(__ldg(pb + 1)) : 0.f;
        } else {
            v[c][0] = Io16<T>::to_float(__ldg(pt));
            v[c][1] = q.in_e ? Io16<T>::to_float(__ldg(pt + 1)) : 0.f;
            v[c][2] = q.in_s ? Io16<T>::to_float(__ldg(pb)) : 0.f;
            v[c][3] = (q.in_e && q.in_s) ? Io16<T>::to_float(__ldg(pb + 1)) : 0.f;
        }
        pt += hw; pb += hw;
    }
}

// grid: (W / blockDim.x, H / (blockDim.y * rows), B), as warp_rows_kernel
template <int MODE, int CT, typename T>
__global__ void __launch_bounds__(256, 5) warp_rows16_kernel(const __grid_constant__ Warp16Args a)
{
    const WarpArgs &g = a.g;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * g.rows;
    if (x >= g.W || y0 >= g.H) return;
    const int y1 = min(y0 + g.rows, g.H);
    const unsigned hw = (unsigned)g.H * g.W, W = (unsigned)g.W;
    const size_t b = blockIdx.z;
    const unsigned short *src = a.img + b * CT * hw;
    asm("" : "+l"(src));     // one opaque 64-bit batch base, 32-bit offsets from it
    const unsigned p0 = (unsigned)y0 * W + (unsigned)x;
    const unsigned short *fl = a.flow + b * 2 * hw + p0;
    unsigned short *out = a.out + b * CT * hw + p0;
    const float xfl = small_int_as_float(x);
    float yfl = small_int_as_float(y0);
    float dx = Io16<T>::to_float(ldg_stream_u16(fl)), dy = Io16<T>::to_float(ldg_stream_u16(fl + hw));
    for (int y = y0; y < y1; ++y) {
        float ndx = 0.f, ndy = 0.f;
        if (y + 1 < y1) {                                                           // next row's flow
            ndx = Io16<T>::to_float(ldg_stream_u16(fl + W));
            ndy = Io16<T>::to_float(ldg_stream_u16(fl + W + hw));
        }
        PixPrep<MODE> cur;
        if constexpr (MODE == FLOWOPS_WARP_GRIDSAMPLE) pix_prep<T>(cur, g, xfl, yfl, x, y, dx, dy);
        else pix_prep(cur, g, xfl, yfl, x, y, dx, dy);
        float v[CT][4];
        pix_gather16<T, CT, MODE>(v, cur, src, hw);
        unsigned short *o = out;
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            stg_stream_u16(o, Io16<T>::from_float(pix_blend(cur, v[c])));
            o += hw;
        }
        dx = ndx; dy = ndy;
        fl += W; out += W; yfl = __fadd_rn(yfl, 1.f);
    }
}

template <int MODE, typename T>
static int launch_warp16(Warp16Args a, cudaStream_t st)
{
    const int B = a.g.B, C = a.g.C;
    const size_t hw = (size_t)a.g.H * a.g.W;
    for (int b0 = 0; b0 < B; b0 += 65535) {            // gridDim.z limit
        Warp16Args c = a;
        c.g.B = B - b0 < 65535 ? B - b0 : 65535;
        c.img = a.img + (size_t)b0 * C * hw; c.flow = a.flow + (size_t)b0 * 2 * hw; c.out = a.out + (size_t)b0 * C * hw;
        dim3 grid, block;
        warp_rows_shape(c.g.B, c.g.H, c.g.W, c.g.rows, grid, block);
        if (C == 3) warp_rows16_kernel<MODE, 3, T><<<grid, block, 0, st>>>(c);
        else if (C == 2) warp_rows16_kernel<MODE, 2, T><<<grid, block, 0, st>>>(c);
        else warp_rows16_kernel<MODE, 1, T><<<grid, block, 0, st>>>(c);
    }
    return check_launch("warp_fwd_16");
}

}  // namespace flowops

using namespace flowops;

extern "C" int flowops_warp_fwd_16(const void *img, const void *flow, void *out, int B, int C, int H, int W, int mode,
                                   const float *lin_x, const float *lin_y, int dtype, void *stream)
{
    FLOWOPS_REQUIRE(img && flow && out, FLOWOPS_EINVAL, "warp_fwd_16: null pointer");
    FLOWOPS_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, FLOWOPS_EINVAL, "warp_fwd_16: bad shape %dx%dx%dx%d", B, C, H, W);
    FLOWOPS_REQUIRE(dtype16_ok(dtype), FLOWOPS_EINVAL, "warp_fwd_16: dtype must be FLOWOPS_DTYPE_F16 or FLOWOPS_DTYPE_BF16, got %d", dtype);
    FLOWOPS_REQUIRE(C <= 3, FLOWOPS_EUNSUPPORTED, "warp_fwd_16: 1..3 channels (frames and flows); cast and call flowops_warp_fwd for %d", C);
    FLOWOPS_REQUIRE((size_t)C * H * W < (1ull << 31), FLOWOPS_EUNSUPPORTED, "warp_fwd_16: C*H*W exceeds int32 indexing");
    FLOWOPS_REQUIRE(H <= (1 << 22) && W <= (1 << 22), FLOWOPS_EUNSUPPORTED, "warp_fwd_16: H, W above 2^22 are not supported");
    FLOWOPS_REQUIRE(mode == FLOWOPS_WARP_RESAMPLE2D || mode == FLOWOPS_WARP_GRIDSAMPLE, FLOWOPS_EINVAL, "warp_fwd_16: unknown mode %d", mode);
    Warp16Args a{};
    a.img = static_cast<const unsigned short *>(img);
    a.flow = static_cast<const unsigned short *>(flow);
    a.out = static_cast<unsigned short *>(out);
    a.g.B = B; a.g.C = C; a.g.H = H; a.g.W = W; a.g.rows = warp_rows_pick(B, H, W);
    a.g.wm1 = (float)(W - 1); a.g.hm1 = (float)(H - 1);
    cudaStream_t st = (cudaStream_t)stream;
    const bool f16 = dtype == FLOWOPS_DTYPE_F16;
    if (mode == FLOWOPS_WARP_RESAMPLE2D)
        return f16 ? launch_warp16<FLOWOPS_WARP_RESAMPLE2D, __half>(a, st)
                   : launch_warp16<FLOWOPS_WARP_RESAMPLE2D, __nv_bfloat16>(a, st);
    FLOWOPS_REQUIRE(lin_x && lin_y, FLOWOPS_EINVAL, "warp_fwd_16: GRIDSAMPLE mode needs the linspace tables");
    FLOWOPS_REQUIRE(H > 1 && W > 1, FLOWOPS_EUNSUPPORTED, "warp_fwd_16: GRIDSAMPLE mode needs H, W > 1");
    float mulx, muly;
    gs_scales(H, W, a.g.invx, a.g.invy, mulx, muly);
    a.g.lin_x = lin_x; a.g.lin_y = lin_y;
    return f16 ? launch_warp16<FLOWOPS_WARP_GRIDSAMPLE, __half>(a, st)
               : launch_warp16<FLOWOPS_WARP_GRIDSAMPLE, __nv_bfloat16>(a, st);
}
