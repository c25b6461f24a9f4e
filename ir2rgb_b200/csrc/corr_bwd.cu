// corr_bwd.cu -- FlowNetC Correlation backward (pad 20, k 1, md 20, s1 1, s2 2) for sm_100a.
//
//   gI1[n,c,y,x] = (1/C) sum_{tj,ti} gO[n,(tj,ti),y,x]        * f2pad[n,c, y+2(tj-10), x+2(ti-10)]
//   gI2[n,c,y,x] = (1/C) sum_{tj,ti} gO[n,(tj,ti),y-2(tj-10),x-2(ti-10)] * f1[n,c,y-2(tj-10),x-2(ti-10)]
//   (reference correlation_cuda_kernel.cu:151-334; terms whose gO coordinate is out of range are dropped)
//
// Both are the same banded contraction   out[c][p] = sum_d G[d][p] * Win[c][p + d]   over the 21 x 21
// displacements d inside one parity plane (see corr_fast.cu):
//   gI1: G = gO,                          Win = f2
//   gI2: G[d'][q] = gO[-d'][q + d'],      Win = f1      (substituting d' = -d; no atomics, deterministic)
// so ONE kernel serves both.  Pre-passes build, in the caller's workspace,
//   * Win as parity planes with channel PAIRS interleaved (c, c+1 adjacent) and shifted right by two
//     columns (TMA start alignment, see corr_fast.cu) -- so that a packed FFMA2 updates two channels;
//   * G as parity planes of gO (gI1) or of the displacement-mirrored, shifted gO (gI2).
// Main kernel, per CTA: 4 plane rows x 32 plane columns x 64 channels, looping over the 21 vertical
// displacements tj.  Per step TMA brings the 21 x 4 x 32 slice of G and ONE new row of Win (the rows
// needed by consecutive tj overlap; a 6-slot ring holds them).  A thread owns 8 pixels x 4 channels
// (16 packed accumulators), keeps the two 28-pair Win windows of its channels in registers and streams
// the 21 x 8 G values: 70 LDS.128 feed 336 FFMA2 per step.  Lanes of a warp differ in channel group,
// so the G loads are broadcasts.
//
// Roofline: FP32-FMA pipe, 2 * (2*B*H*W*441*C) FLOP for both gradients (dense count).  This contraction
// has less register-level reuse than the forward (each G value feeds only the thread's 4 channels), so
// shared-memory bandwidth, not the FMA pipe, is the first limit -- see DESIGN.md.
#include "corr.cuh"
#include "tma.cuh"

namespace flowops {

constexpr int bD = 21, bR = 10, bShift = 2;
constexpr int bTY = 4, bTX = 32, bPX = 8;
constexpr int bCP = 32;                                  // channel pairs per CTA (64 channels)
constexpr int bGFloats = bD * bTY * bTX;                 // 2688: one tj slice of G
constexpr int bWRowPairs = 54;                           // 32 + 20 window pairs + 2 (bank spreading)
constexpr int bWRowFloats = bCP * bWRowPairs * 2;        // 3456: one Win row for 32 channel pairs
constexpr int bNG = 3, bNW = bTY + 2;                    // G stages, Win ring slots
constexpr int bSmemBytes = (bNG * bGFloats + bNW * bWRowFloats) * 4;   // 115200
constexpr int bEpiPitch = 36;
static_assert((bGFloats * 4) % 128 == 0 && (bWRowFloats * 4) % 128 == 0, "TMA destinations stay 128-byte aligned");
static_assert(2 * bCP * bTY * bEpiPitch * 4 <= bNW * bWRowFloats * 4, "epilogue staging fits in the Win ring");

struct BwdGeom {
    int Hp, Wp, pitch1, pitch2, Cp2;     // plane rows/cols, G row pitch, Win row pitch (pairs), channel pairs
    size_t win_floats, g_floats;
};

static inline BwdGeom bwd_geom(const CorrGeom &g)
{
    BwdGeom b;
    b.Hp = (g.H + 1) / 2;
    b.Wp = (g.W + 1) / 2;
    b.pitch1 = (b.Wp + 3) & ~3;
    b.pitch2 = (b.Wp + bShift + 3) & ~3;
    b.Cp2 = (g.C + 1) / 2;
    b.win_floats = (size_t)g.B * 4 * b.Cp2 * b.Hp * b.pitch2 * 2;
    b.g_floats = (size_t)g.B * 4 * (bD * bD) * b.Hp * b.pitch1;
    return b;
}

size_t corr_fast_bwd_workspace(const CorrGeom &g)
{
    const BwdGeom b = bwd_geom(g);
    return (b.win_floats + b.g_floats) * sizeof(float);
}

// ---------------------------------------------------------------------------------------------
// pre-passes
// ---------------------------------------------------------------------------------------------

// NCHW -> Win[n*4+plane][c/2][Hp][pitch2][2]  (channel pairs interleaved, shifted by bShift columns)
__global__ void __launch_bounds__(256) corr_bwd_pairs(const float *__restrict__ in, float *__restrict__ WP,
                                                      int B, int C, int H, int W, int Hp, int pitch2, int Cp2, int vec_ok)
{
    const int W4 = (W + 3) >> 2;
    const size_t hw = (size_t)H * W;
    const size_t total = (size_t)B * Cp2 * H * W4;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(idx % W4) * 4;
        size_t r = idx / W4;
        const int y = (int)(r % H); r /= H;
        const int cp = (int)(r % Cp2);
        const int n = (int)(r / Cp2);
        const float *s0 = in + ((size_t)n * C + 2 * cp) * hw + (size_t)y * W + x;
        const bool has1 = 2 * cp + 1 < C;
        float4 v0, v1 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (vec_ok) {
            v0 = ldg_stream4(s0);
            if (has1) v1 = ldg_stream4(s0 + hw);
        } else {
            v0.x = s0[0]; v0.y = x + 1 < W ? s0[1] : 0.f; v0.z = x + 2 < W ? s0[2] : 0.f; v0.w = x + 3 < W ? s0[3] : 0.f;
            if (has1) { v1.x = s0[hw]; v1.y = x + 1 < W ? s0[hw + 1] : 0.f; v1.z = x + 2 < W ? s0[hw + 2] : 0.f; v1.w = x + 3 < W ? s0[hw + 3] : 0.f; }
        }
        const int py = y & 1, yy = y >> 1, xx = (x >> 1) + bShift;
        float *row0 = WP + ((((size_t)n * 4 + py * 2 + 0) * Cp2 + cp) * Hp + yy) * (size_t)(pitch2 * 2);
        float *row1 = WP + ((((size_t)n * 4 + py * 2 + 1) * Cp2 + cp) * Hp + yy) * (size_t)(pitch2 * 2);
        // two plane columns x two channels = 16 aligned bytes per parity
        *reinterpret_cast<float4 *>(row0 + 2 * xx) = make_float4(v0.x, v1.x, v0.z, v1.z);
        *reinterpret_cast<float4 *>(row1 + 2 * xx) = make_float4(v0.y, v1.y, v0.w, v1.w);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (x == 0) { *reinterpret_cast<float4 *>(row0) = z; *reinterpret_cast<float4 *>(row1) = z; }
        if (x + 4 >= W)
            for (int q = xx + 2; q < pitch2; q += 2) {
                *reinterpret_cast<float4 *>(row0 + 2 * q) = z;
                *reinterpret_cast<float4 *>(row1 + 2 * q) = z;
            }
    }
}

// gO [n][441][H][W] -> G[n*4+plane][441][Hp][pitch1]; mirrored = 0: plain parity planes (gI1);
// mirrored = 1: G[d'][y][x] = gO[440 - d'][y + tj'][x + ti'] in plane coordinates, zero outside (gI2).
__global__ void __launch_bounds__(256) corr_bwd_gplanes(const float *__restrict__ gout, float *__restrict__ G,
                                                        int B, int H, int W, int Hp, int pitch1, int mirrored)
{
    const int P4 = pitch1 >> 2;
    const size_t total = (size_t)B * 4 * (bD * bD) * Hp * P4;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(idx % P4) * 4;
        size_t r = idx / P4;
        const int y = (int)(r % Hp); r /= Hp;
        const int d = (int)(r % (bD * bD)); r /= (bD * bD);
        const int plane = (int)(r & 3);
        const int n = (int)(r >> 2);
        const int py = plane >> 1, px = plane & 1;
        int src_d = d, ys = y, xs = x;
        if (mirrored) {
            src_d = bD * bD - 1 - d;
            ys = y + (d / bD - bR);
            xs = x + (d % bD - bR);
        }
        const int Y = 2 * ys + py;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (Y >= 0 && Y < H) {
            const float *src = gout + (((size_t)n * (bD * bD) + src_d) * H + Y) * W;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int X = 2 * (xs + q) + px;
                if (X >= 0 && X < W) v[q] = __ldg(src + X);
            }
        }
        *reinterpret_cast<float4 *>(G + idx * 4) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
corr_bwd_tile(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmW,
              float *__restrict__ out, int C, int H, int W, int row_tiles, int x_tiles, int c_chunks)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bar[bNG];

    int bid = blockIdx.x;
    const int cc = bid % c_chunks; bid /= c_chunks;      // channel chunk fastest: neighbours share the G slice in L2
    const int plane = bid & 3; bid >>= 2;
    const int xt = bid % x_tiles; bid /= x_tiles;
    const int rt = bid % row_tiles;
    const int n = bid / row_tiles;
    const int py = plane >> 1, px = plane & 1;
    const int y0 = rt * bTY, x0 = xt * bTX;
    const int np = n * 4 + plane;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cg = tid & 15;                 // channel group: pairs cg and cg + 16 of the chunk
    const int pg = tid >> 4;                 // pixel group
    const int xb = pg & 3, yi = pg >> 2;

    float *gbuf = reinterpret_cast<float *>(smem);
    float *wbuf = gbuf + bNG * bGFloats;
    const uint32_t g_base = smem_u32(gbuf), w_base = smem_u32(wbuf), bar_base = smem_u32(full_bar);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < bNG; ++s) mbar_init(bar_base + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // step t needs G slice t and Win rows t .. t+bTY-1 (relative to y0 - bR); row t+bTY-1 is the new one
    auto issue = [&](int t) {
        const uint32_t bar = bar_base + 8 * (t % bNG);
        const int first_row = t == 0 ? 0 : t + bTY - 1;
        const int n_rows = t == 0 ? bTY : 1;
        mbar_expect_tx(bar, (bGFloats + n_rows * bWRowFloats) * 4);
        tma_load_4d(g_base + (t % bNG) * bGFloats * 4, &tmG, x0, y0, t * bD, np, bar);
        for (int rr = first_row; rr < first_row + n_rows; ++rr)
            tma_load_4d(w_base + (rr % bNW) * bWRowFloats * 4, &tmW, 2 * (x0 - bR + bShift), y0 - bR + rr, cc * bCP, np, bar);
    };
    if (tid == 0) { issue(0); issue(1); }

    float2 acc[2][bPX];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int k = 0; k < bPX; ++k) acc[h][k] = make_float2(0.f, 0.f);

    for (int t = 0; t < bD; ++t) {
        mbar_wait(bar_base + 8 * (t % bNG), (t / bNG) & 1);
        const float *wrow = wbuf + ((t + yi) % bNW) * bWRowFloats + 2 * (xb * bPX);
        const float *gsl = gbuf + (t % bNG) * bGFloats + yi * bTX + xb * bPX;
        float2 w2[2][bPX + bD - 1];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float4 *pw = reinterpret_cast<const float4 *>(wrow + (cg + 16 * h) * (bWRowPairs * 2));
#pragma unroll
            for (int q = 0; q < (bPX + bD - 1) / 2; ++q) {
                const float4 v = pw[q];
                w2[h][2 * q] = make_float2(v.x, v.y); w2[h][2 * q + 1] = make_float2(v.z, v.w);
            }
        }
#pragma unroll
        for (int i = 0; i < bD; ++i) {
            const float4 g0 = *reinterpret_cast<const float4 *>(gsl + i * (bTY * bTX));
            const float4 g1 = *reinterpret_cast<const float4 *>(gsl + i * (bTY * bTX) + 4);
            const float g[bPX] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int k = 0; k < bPX; ++k) acc[h][k] = fma2(make_float2(g[k], g[k]), w2[h][k + i], acc[h][k]);
        }
        __syncthreads();
        if (tid == 0 && t + 2 < bD) issue(t + 2);
    }

    // ---- epilogue: scale, transpose through shared memory, strided store into NCHW ----
    float *stage = wbuf;
    const float nelems = (float)C, inv_nelems = 1.0f / nelems;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int cl = 2 * (cg + 16 * h);            // channel inside the chunk (x half), +1 (y half)
        float lo[bPX], hi[bPX];
#pragma unroll
        for (int k = 0; k < bPX; ++k) {
            lo[k] = div_nelems(acc[h][k].x, nelems, inv_nelems);
            hi[k] = div_nelems(acc[h][k].y, nelems, inv_nelems);
        }
        float *d0 = stage + ((cl)*bTY + yi) * bEpiPitch + xb * bPX;
        float *d1 = stage + ((cl + 1) * bTY + yi) * bEpiPitch + xb * bPX;
        *reinterpret_cast<float4 *>(d0) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<float4 *>(d0 + 4) = make_float4(lo[4], lo[5], lo[6], lo[7]);
        *reinterpret_cast<float4 *>(d1) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4 *>(d1 + 4) = make_float4(hi[4], hi[5], hi[6], hi[7]);
    }
    __syncthreads();
    const int x = 2 * (x0 + lane) + px;
    const size_t hw = (size_t)H * W;
    for (int row = warp; row < 2 * bCP * bTY; row += 8) {
        const int cl = row / bTY, ry = row - cl * bTY;
        const int c = cc * (2 * bCP) + cl;
        const int y = 2 * (y0 + ry) + py;
        if (c < C && y < H && x < W)
            out[((size_t)n * C + c) * hw + (size_t)y * W + x] = stage[row * bEpiPitch + lane];
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static inline unsigned prepass_grid(size_t total)
{
    size_t blocks = (total + 255) / 256;
    if (blocks > (size_t)kNumSMs * 8 * 8) blocks = (size_t)kNumSMs * 8 * 8;
    return (unsigned)(blocks < 1 ? 1 : blocks);
}

int corr_fast_bwd_launch(const float *in1, const float *in2, const float *gout, float *gin1, float *gin2,
                         const CorrGeom &g, void *ws, size_t ws_bytes, cudaStream_t st)
{
    const BwdGeom b = bwd_geom(g);
    const size_t need = corr_fast_bwd_workspace(g);
    FLOWOPS_REQUIRE(ws && ws_bytes >= need && ((uintptr_t)ws & 255) == 0, FLOWOPS_EWORKSPACE,
                    "corr_bwd: workspace of %zu bytes (256-byte aligned) required, got %zu", need, ws_bytes);
    float *WP = reinterpret_cast<float *>(ws);
    float *G = WP + b.win_floats;

    {   // per-device attribute: set before every launch (see corr_fast.cu), not once per process
        const cudaError_t e = cudaFuncSetAttribute(corr_bwd_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, bSmemBytes);
        if (e != cudaSuccess) { set_error("corr_bwd: cannot reserve %d bytes of shared memory: %s", bSmemBytes, cudaGetErrorString(e)); return (int)e; }
    }

    CUtensorMap tmG, tmW;
    {
        const cuuint64_t row = (cuuint64_t)b.pitch1 * 4, img = row * b.Hp;
        const cuuint64_t dims[4] = {(cuuint64_t)b.pitch1, (cuuint64_t)b.Hp, (cuuint64_t)(bD * bD), (cuuint64_t)g.B * 4};
        const cuuint64_t strides[3] = {row, img, img * (bD * bD)};
        const cuuint32_t box[4] = {bTX, bTY, bD, 1};
        const int rc = encode_map4(&tmG, G, dims, strides, box, "corr_bwd");
        if (rc) return rc;
    }
    {
        const cuuint64_t row = (cuuint64_t)b.pitch2 * 8, img = row * b.Hp;
        const cuuint64_t dims[4] = {(cuuint64_t)b.pitch2 * 2, (cuuint64_t)b.Hp, (cuuint64_t)b.Cp2, (cuuint64_t)g.B * 4};
        const cuuint64_t strides[3] = {row, img, img * b.Cp2};
        const cuuint32_t box[4] = {bWRowPairs * 2, 1, bCP, 1};
        const int rc = encode_map4(&tmW, WP, dims, strides, box, "corr_bwd");
        if (rc) return rc;
    }

    const int row_tiles = (b.Hp + bTY - 1) / bTY, x_tiles = (b.Wp + bTX - 1) / bTX;
    const int c_chunks = (b.Cp2 + bCP - 1) / bCP;
    const size_t grid = (size_t)g.B * 4 * row_tiles * x_tiles * c_chunks;
    FLOWOPS_REQUIRE(grid < (1ull << 31), FLOWOPS_EUNSUPPORTED, "corr_bwd: grid too large");
    const int vec_ok = (g.W % 4 == 0) && aligned16(in1) && aligned16(in2);
    const bool padded = (g.W & 7) || (g.H & 1);      // planes then contain cells no pre-pass thread writes

    for (int which = 0; which < 2; ++which) {
        float *dst = which == 0 ? gin1 : gin2;
        if (!dst) continue;
        const float *win_src = which == 0 ? in2 : in1;
        if (padded) {
            cudaError_t e = cudaMemsetAsync(WP, 0, b.win_floats * sizeof(float), st);
            if (e != cudaSuccess) { set_error("corr_bwd: memset failed: %s", cudaGetErrorString(e)); return (int)e; }
        }
        corr_bwd_pairs<<<prepass_grid((size_t)g.B * b.Cp2 * g.H * ((g.W + 3) / 4)), 256, 0, st>>>(
            win_src, WP, g.B, g.C, g.H, g.W, b.Hp, b.pitch2, b.Cp2, vec_ok);
        corr_bwd_gplanes<<<prepass_grid(b.g_floats / 4), 256, 0, st>>>(gout, G, g.B, g.H, g.W, b.Hp, b.pitch1, which);
        corr_bwd_tile<<<(unsigned)grid, 256, bSmemBytes, st>>>(tmG, tmW, dst, g.C, g.H, g.W, row_tiles, x_tiles, c_chunks);
        const int rc = check_launch(which == 0 ? "corr_bwd(grad input1)" : "corr_bwd(grad input2)");
        if (rc) return rc;
    }
    return 0;
}

}  // namespace flowops
