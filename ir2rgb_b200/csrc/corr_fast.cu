// corr_fast.cu -- FlowNetC Correlation (pad 20, k 1, md 20, s1 1, s2 2 -> 21x21 = 441 channels)
// for sm_100a: TMA-staged, register-tiled FP32-FMA kernel.
//
//   out[n, tj*21+ti, y, x] = (1/C) * sum_c f1[n,c,y,x] * f2pad[n,c, y + 2(tj-10), x + 2(ti-10)]
//   (reference correlation_cuda_kernel.cu:74-147; tj = vertical = slow index)
//
// Structure
//   stride2 = 2 means a pixel only ever meets pixels of its own (row, column) parity, so the
//   problem splits into four independent "parity planes" in which the displacement is a dense
//   +-10 x +-10 window.  A pre-pass (corr_planarize) de-interleaves both inputs into planes
//   P[n][plane][c][H/2][W/2] in the caller-provided workspace (this replaces the reference's
//   zero-padded NHWC scratch copies rInput1/rInput2 -- same size, no padding: TMA's out-of-bounds
//   zero fill supplies it).  The main kernel then runs, per CTA, a tile of 3 plane-rows x 32 plane-
//   columns x all 441 displacements:
//     * per pipeline stage TMA brings 8 channels of the f1 tile (3 x 36) and of the displaced f2
//       window (23 x 52) into shared memory, 4 stages deep, completion on mbarriers;
//     * 252 threads each own a register tile of 8 pixels x 21 horizontal displacements for one
//       (row, vertical displacement) pair: per channel 2 + 7 LDS.128 feed 168 FFMAs
//       (sliding-window reuse of the f2 row inside registers);
//     * a warp holds 32 (row, tj) pairs that touch only 13 distinct f2 rows and 3 f1 rows, so its
//       shared-memory loads are mostly broadcasts (about 16 wavefronts per 168 FFMA issue slots);
//     * the epilogue transposes the accumulators through shared memory so that every global store
//       instruction writes one 32-pixel output row segment.
//
// Roofline: FP32-FMA pipe.  Algorithmic work 2*B*H*W*441*C FLOP (dense count, taps that fall in
// the zero padding included -- the kernel does not skip them).
#include "corr.cuh"
#include "io16.cuh"
#include "tma.cuh"

namespace flowops {

constexpr int kD = 21;                       // displacements per axis
constexpr int kR = 10;                       // displacement radius in plane coordinates
constexpr int kTY = 3;                       // plane rows per CTA tile
constexpr int kTX = 32;                      // plane columns per CTA tile
constexpr int kPX = 8;                       // pixels per thread
constexpr int kCK = 8;                       // channels per pipeline stage
constexpr int kStages = 4;
constexpr int kF2W = kTX + 2 * kR;           // 52
constexpr int kF2H = kTY + 2 * kR;           // 23
constexpr int kF1W = kTX + 4;                // 36: +4 columns so the 3 f1 rows sit in different banks
constexpr int kF2Floats = kCK * kF2H * kF2W; // 9568
constexpr int kF1Floats = kCK * kTY * kF1W;  // 864
constexpr int kStageBytes = (kF2Floats + kF1Floats) * 4;   // 41728 (multiple of 128)
constexpr int kPairs = kTY * kD;             // 63 (row, tj) pairs per tile column block
constexpr int kEpiPitch = 36;                // floats per staged output row
constexpr int kEpiGroup = 11;                // tj values staged per epilogue pass: two passes (11 + 10)
constexpr int kEpiRows = kEpiGroup * kD * kTY;              // 693
constexpr int kSmemBytes = kStages * kStageBytes;           // 166912
static_assert(kEpiRows * kEpiPitch * 4 <= kSmemBytes, "epilogue staging must fit in the pipeline buffers");
static_assert(kStageBytes % 128 == 0 && (kF2Floats * 4) % 128 == 0, "TMA destinations must stay 128-byte aligned");

// TMA needs the innermost start coordinate of a box to be 16-byte aligned (measured on B200: a box
// starting at x = -10 or 2 raises "illegal instruction", -8 / 0 / -12 are fine).  The f2 window starts
// kR = 10 columns left of the tile, so the f2 planes are stored shifted right by kShift = 2 columns
// (two explicit zero columns in front): the window then starts at x0 - 8.
constexpr int kShift = 2;

struct PlaneGeom {
    int Hp, Wp;             // plane rows / valid plane columns (ceil)
    int pitch1, pitch2;     // row pitch in floats of the f1 / f2 planes (multiples of 4)
    size_t elems1, elems2;  // floats per (n, plane, c) image
};

static inline PlaneGeom plane_geom(const CorrGeom &g)
{
    PlaneGeom p;
    p.Hp = (g.H + 1) / 2;
    p.Wp = (g.W + 1) / 2;
    p.pitch1 = (p.Wp + 3) & ~3;
    p.pitch2 = (p.Wp + kShift + 3) & ~3;
    p.elems1 = (size_t)p.Hp * p.pitch1;
    p.elems2 = (size_t)p.Hp * p.pitch2;
    return p;
}

bool corr_fast_supported(const CorrGeom &g)
{
    return g.k == 1 && g.s1 == 1 && g.s2 == 2 && g.pad == g.md && g.md == 2 * kR && g.D == kD;
}

size_t corr_fast_fwd_workspace(const CorrGeom &g)
{
    const PlaneGeom p = plane_geom(g);
    return sizeof(float) * (size_t)g.B * 4 * g.C * (p.elems1 + p.elems2);
}

// ---------------------------------------------------------------------------------------------
// pre-pass: NCHW -> parity planes  P[n][py*2+px][c][Hp][pitch]
// ---------------------------------------------------------------------------------------------
// TI = float: the inputs as the reference takes them.  TI = __half / __nv_bfloat16: 16-bit features widened on the
// way in (the `.float()` of FlowNetC.py:86-87 folded into this pass; exact, so the planes hold the same values).
template <typename TI>
__device__ __forceinline__ float4 planarize_load4(const TI *src, int x, int W, int vec_ok)
{
    float4 v;
    if constexpr (std::is_same<TI, float>::value) {
        if (vec_ok) {
            v = ldg_stream4(src);
        } else {
            v.x = src[0];
            v.y = x + 1 < W ? src[1] : 0.f;
            v.z = x + 2 < W ? src[2] : 0.f;
            v.w = x + 3 < W ? src[3] : 0.f;
        }
    } else {
        const unsigned short *s16 = reinterpret_cast<const unsigned short *>(src);
        if (vec_ok) {                                  // four 16-bit elements: one 8-byte load
            uint2 w;
            asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(w.x), "=r"(w.y) : "l"(s16));
            v.x = Io16<TI>::to_float((unsigned short)(w.x & 0xffffu)); v.y = Io16<TI>::to_float((unsigned short)(w.x >> 16));
            v.z = Io16<TI>::to_float((unsigned short)(w.y & 0xffffu)); v.w = Io16<TI>::to_float((unsigned short)(w.y >> 16));
        } else {
            v.x = Io16<TI>::to_float(s16[0]);
            v.y = x + 1 < W ? Io16<TI>::to_float(s16[1]) : 0.f;
            v.z = x + 2 < W ? Io16<TI>::to_float(s16[2]) : 0.f;
            v.w = x + 3 < W ? Io16<TI>::to_float(s16[3]) : 0.f;
        }
    }
    return v;
}

template <typename TI>
__global__ void __launch_bounds__(256) corr_planarize(const TI *__restrict__ in1, const TI *__restrict__ in2,
                                                      float *__restrict__ P1, float *__restrict__ P2,
                                                      int B, int C, int H, int W, int Hp, int pitch1, int pitch2, int vec_ok)
{
    const bool second = blockIdx.y != 0;
    const TI *__restrict__ in = second ? in2 : in1;
    float *__restrict__ P = second ? P2 : P1;
    const int pitch = second ? pitch2 : pitch1;
    const int shift = second ? kShift : 0;
    const int W4 = (W + 3) >> 2;
    const size_t total = (size_t)B * C * H * W4;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(idx % W4) * 4;
        size_t r = idx / W4;
        const int y = (int)(r % H); r /= H;
        const int c = (int)(r % C);
        const int n = (int)(r / C);
        const TI *src = in + (((size_t)n * C + c) * H + y) * W + x;
        const float4 v = planarize_load4<TI>(src, x, W, vec_ok);
        const int py = y & 1, yy = y >> 1, xx = (x >> 1) + shift;
        float *row0 = P + ((((size_t)n * 4 + py * 2 + 0) * C + c) * Hp + yy) * pitch;
        float *row1 = P + ((((size_t)n * 4 + py * 2 + 1) * C + c) * Hp + yy) * pitch;
        // pitch is a multiple of 4 and xx is even: 8-byte aligned pairs
        *reinterpret_cast<float2 *>(row0 + xx) = make_float2(v.x, v.z);
        *reinterpret_cast<float2 *>(row1 + xx) = make_float2(v.y, v.w);
        if (second) {   // explicit zero columns in front of and behind the shifted row
            if (x == 0) {
                *reinterpret_cast<float2 *>(row0) = make_float2(0.f, 0.f);
                *reinterpret_cast<float2 *>(row1) = make_float2(0.f, 0.f);
            }
            if (x + 4 >= W)
                for (int q = xx + 2; q < pitch; q += 2) {
                    *reinterpret_cast<float2 *>(row0 + q) = make_float2(0.f, 0.f);
                    *reinterpret_cast<float2 *>(row1 + q) = make_float2(0.f, 0.f);
                }
        }
    }
}

// Same planes from channels-last (NHWC) inputs -- what a channels_last conv body hands over (FlowNetC's
// conv3 features): one pass instead of an NHWC->NCHW copy followed by corr_planarize.  A CTA transposes
// 64 pixels x 32 channels of one image row through shared memory: loads are coalesced along c (128 B per
// pixel), stores along the plane row (128 B per channel and parity).
constexpr int kNhwcPix = 64, kNhwcCh = 32;
//
// With `bias` the kernel is also the epilogue of the convolution that produced the features (FlowNetC's conv3,
// FlowNetC.py:23,75-81): t = y + bias[c]; v = t > 0 ? t : t * slope is applied on the fly, the planes receive v,
// and v is written back in NHWC order only where another consumer needs it (`act`; frame 1 feeds conv_redir,
// frame 2 feeds nothing else).  That replaces  bias+LeakyReLU (read+write)  +  planarize (read+write)  by one read
// and one or two writes, and leaves the correlation proper with no pre-pass of its own.
// `act` may be the input itself (in-place epilogue, CorrelationPlanes.fill_from_conv_(write_act=True)): in1 / in2 / act
// are therefore not __restrict__ and the feature loads are plain ld.global, not the non-coherent .nc form, which PTX
// leaves undefined for memory the kernel also writes.
__global__ void __launch_bounds__(256) corr_planarize_nhwc(const float *in1, const float *in2,
                                                           float *__restrict__ P1, float *__restrict__ P2,
                                                           int C, int H, int W, int Hp, int Wp, int pitch1, int pitch2,
                                                           int only, const float *__restrict__ bias, float slope,
                                                           float *act)
{
    __shared__ float tile[2][kNhwcPix / 2][kNhwcCh + 1];       // [column parity][plane column][channel]
    const bool second = only < 0 ? (blockIdx.z & 1) : (only == 1);
    const int n = only < 0 ? (blockIdx.z >> 1) : blockIdx.z;
    const float *in = second ? in2 : in1;
    float *__restrict__ P = second ? P2 : P1;
    const int pitch = second ? pitch2 : pitch1, shift = second ? kShift : 0;
    const int c_tiles = (C + kNhwcCh - 1) / kNhwcCh;
    const int ct = blockIdx.x % c_tiles, xt = blockIdx.x / c_tiles;
    const int y = blockIdx.y, c0 = ct * kNhwcCh, x0 = xt * kNhwcPix;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    const size_t row_off = ((size_t)n * H + y) * W * C;
    const float *row = in + row_off;
    const float b = (bias && c0 + lane < C) ? __ldg(bias + c0 + lane) : 0.f;
#pragma unroll
    for (int i = 0; i < kNhwcPix / 8; ++i) {
        const int px = warp + 8 * i, x = x0 + px, c = c0 + lane;
        float v = 0.f;
        if (x < W && c < C) {
            v = row[(size_t)x * C + c];
            if (bias) {
                const float t = __fadd_rn(v, b);
                v = t > 0.f ? t : __fmul_rn(t, slope);
                if (act) act[row_off + (size_t)x * C + c] = v;
            }
        }
        tile[px & 1][px >> 1][lane] = v;
    }
    __syncthreads();
    const int py = y & 1, yy = y >> 1;
    const int xx0 = x0 >> 1;                                    // first plane column of this tile
#pragma unroll
    for (int i = 0; i < 2 * kNhwcCh / 8; ++i) {
        const int r = warp + 8 * i;                             // (parity, channel) row of the tile
        const int par = r / kNhwcCh, cl = r - par * kNhwcCh, c = c0 + cl;
        if (c >= C) continue;
        float *dst = P + ((((size_t)n * 4 + py * 2 + par) * C + c) * Hp + yy) * pitch;
        const int xx = xx0 + lane;
        if (xx + shift < pitch) dst[xx + shift] = tile[par][lane][cl];     // columns beyond Wp hold zeros already
        if (shift && xt == 0 && lane < shift) dst[lane] = 0.f;             // leading zero columns of the f2 planes
        if (xx0 + kNhwcPix / 2 >= Wp)                                      // last tile: zero the tail of the pitch
            for (int q = Wp + shift + lane; q < pitch; q += 32)
                if (q >= xx0 + kNhwcPix / 2 + shift) dst[q] = 0.f;
    }
}

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
// channels-last epilogue staging: [row of the tile][pixel][channel of the group], pitches chosen so that the 21
// (tj, row) writers of a group hit 21 different banks (see the epilogue)
constexpr int kNhwcGroup = 11;                                // tj values per epilogue pass: two passes (11 + 10)
constexpr int kNhwcTjPitch = 24;                              // the 21 channels of one tj, padded so that they start 16-byte aligned
constexpr int kNhwcPixPitch = kNhwcGroup * kNhwcTjPitch + 4;  // 268 floats per staged pixel
constexpr int kNhwcRowPitch = kTX * kNhwcPixPitch + 8;        // 8584
static_assert(kTY * kNhwcRowPitch * 4 <= kSmemBytes, "NHWC epilogue staging must fit in the pipeline buffers");

// NHWC_OUT: the cost volume is written channels-last into channels [c_off, c_off + 441) of a [B, H, W, c_dst]
// tensor, with LeakyReLU(slope) applied (slope 1 = identity) -- FlowNetC's `corr_activation` and the concat with
// conv_redir (FlowNetC.py:89-94) folded into the store, for a channels_last conv body.
// OT: storage type of the (NCHW) output; 16-bit types round on the way out (the `.half()` of FlowNetC.py:87).
template <int UNROLL, bool NHWC_OUT, typename OT = float>
__global__ void __launch_bounds__(256, 1)
corr_fwd_fast(const __grid_constant__ CUtensorMap tm1, const __grid_constant__ CUtensorMap tm2,
              float *__restrict__ out, int C, int H, int W, int row_tiles, int x_tiles,
              int c_dst, int c_off, float slope)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bar[kStages];
    __shared__ __align__(8) uint64_t empty_bar[kStages];       // one arrival per warp when it is done with a slot

    // tile decode: plane fastest so the two column parities of a row segment are written close in time
    int bid = blockIdx.x;
    const int plane = bid & 3; bid >>= 2;
    const int xt = bid % x_tiles; bid /= x_tiles;
    const int rt = bid % row_tiles;
    const int n = bid / row_tiles;
    const int py = plane >> 1, px = plane & 1;
    const int y0 = rt * kTY, x0 = xt * kTX;          // plane coordinates of the tile origin
    const int np = n * 4 + plane;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int xb = warp >> 1;                         // 8-pixel column block 0..3
    const int pair = (warp & 1) * 32 + lane;          // 0..63, 63 is a spare lane
    const bool live = pair < kPairs;
    const int pp = live ? pair : kPairs - 1;
    const int tj = pp / kTY, yi = pp - tj * kTY;

    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_u32(full_bar);
    const uint32_t empty_base = smem_u32(empty_bar);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) { mbar_init(bar_base + 8 * s, 1); mbar_init(empty_base + 8 * s, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int n_it = (C + kCK - 1) / kCK;
    auto issue = [&](int it, int slot) {
        const uint32_t bar = bar_base + 8 * slot;
        const uint32_t dst = smem_base + slot * kStageBytes;
        mbar_expect_tx(bar, kStageBytes);
        tma_load_4d(dst, &tm2, x0 - kR + kShift, y0 - kR, it * kCK, np, bar);   // x0 - 8: 16-byte aligned
        tma_load_4d(dst + kF2Floats * 4, &tm1, x0, y0, it * kCK, np, bar);
    };
    if (tid == 0) {
        for (int s = 0; s < kStages && s < n_it; ++s) issue(s, s);
    }

    // Accumulators.  The FP32 pipe only sustains its rate when register-file traffic is low: a scalar
    // FFMA with three distinct source registers runs this sliding-window pattern at 54 % of peak on B200
    // (register-bank conflicts), the packed fma.rn.f32x2 form at 87 % (tools/ffma_probe.cu).  So the
    // 21 x 8 tile is held as 64-bit pairs over the displacement index i, paired where i + k is even so
    // that (w[i+k], w[i+k+1]) is an aligned register pair straight out of an LDS.128:
    //   even k: pairs (i = 2p, 2p+1), p = 0..9, single i = 20;   odd k: single i = 0, pairs (i = 2p+1, 2p+2)
    float2 accp[kPX][10];
    float accs[kPX];
#pragma unroll
    for (int k = 0; k < kPX; ++k) {
        accs[k] = 0.f;
#pragma unroll
        for (int p = 0; p < 10; ++p) accp[k][p] = make_float2(0.f, 0.f);
    }

    const int f2_off = (yi + tj) * kF2W + xb * kPX;
    const int f1_off = yi * kF1W + xb * kPX;

    for (int it = 0; it < n_it; ++it) {
        const int slot = it % kStages;
        mbar_wait(bar_base + 8 * slot, (it / kStages) & 1);
        const float *f2s = reinterpret_cast<const float *>(smem + slot * kStageBytes) + f2_off;
        const float *f1s = reinterpret_cast<const float *>(smem + slot * kStageBytes) + kF2Floats + f1_off;
#pragma unroll UNROLL
        for (int ck = 0; ck < kCK; ++ck) {
            float a[kPX];
            float2 w2[(kPX + kD - 1) / 2];       // w2[j] = (w[2j], w[2j+1])
            const float4 *pa = reinterpret_cast<const float4 *>(f1s + ck * (kTY * kF1W));
            const float4 *pw = reinterpret_cast<const float4 *>(f2s + ck * (kF2H * kF2W));
#pragma unroll
            for (int q = 0; q < kPX / 4; ++q) {
                const float4 v = pa[q];
                a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
            }
#pragma unroll
            for (int q = 0; q < (kPX + kD - 1) / 4; ++q) {
                const float4 v = pw[q];
                w2[2 * q] = make_float2(v.x, v.y); w2[2 * q + 1] = make_float2(v.z, v.w);
            }
#pragma unroll
            for (int k = 0; k < kPX; ++k) {
                const float2 ad = make_float2(a[k], a[k]);
#pragma unroll
                for (int p = 0; p < 10; ++p) {
                    const int j = (k & 1) ? (2 * p + 1 + k) / 2 : (2 * p + k) / 2;
                    accp[k][p] = fma2(ad, w2[j], accp[k][p]);
                }
                accs[k] = (k & 1) ? __fmaf_rn(a[k], w2[k / 2].y, accs[k])
                                  : __fmaf_rn(a[k], w2[(kD - 1 + k) / 2].x, accs[k]);
            }
        }
        // Slot recycling without a CTA-wide barrier: each warp signals the slot's "empty" mbarrier, and the
        // issuing thread refills the slot consumed one iteration EARLIER (by now every warp has normally left
        // it), so the warps are free to drift apart and keep the FMA pipe fed while one of them loads.
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_base + 8 * slot);
        if (tid == 0 && it >= 1 && it - 1 + kStages < n_it) {
            const int prev = (it - 1) % kStages;
            mbar_wait(empty_base + 8 * prev, ((it - 1) / kStages) & 1);
            issue(it - 1 + kStages, prev);
        }
    }
    __syncthreads();                                  // all warps are out of the pipeline buffers

    // ---- epilogue: registers -> shared (row-major 32-pixel segments) -> global ----
    // (the last loop iteration ended with __syncthreads and no TMA is in flight)
    float *stage = reinterpret_cast<float *>(smem);
    const float nelems = (float)C;
    const float inv_nelems = 1.0f / nelems;
    const size_t hw = (size_t)H * W;
    float *out_n = out + (size_t)n * (kD * kD) * hw;
    if constexpr (NHWC_OUT) {
#pragma unroll 1
        for (int grp = 0; grp < (kD + kNhwcGroup - 1) / kNhwcGroup; ++grp) {
            const int tj0 = grp * kNhwcGroup;
            const int n_ch = (min(kD, tj0 + kNhwcGroup) - tj0) * kD;          // 231, then 210
            if (live && tj >= tj0 && tj < tj0 + kNhwcGroup) {
                // stage[row][pixel][tj-in-group][24]: a thread's 21 channels of one pixel are contiguous and 16-byte
                // aligned -> 5 STS.128 + 1 STS.32 per pixel
                float *dst = stage + yi * kNhwcRowPitch + (xb * kPX) * kNhwcPixPitch + (tj - tj0) * kNhwcTjPitch;
#pragma unroll
                for (int k = 0; k < kPX; ++k) {
                    float v[kD];
#pragma unroll
                    for (int i = 0; i < kD; ++i) {
                        float t;
                        if (k & 1) t = (i == 0) ? accs[k] : (((i - 1) & 1) ? accp[k][(i - 1) / 2].y : accp[k][(i - 1) / 2].x);
                        else       t = (i == kD - 1) ? accs[k] : ((i & 1) ? accp[k][i / 2].y : accp[k][i / 2].x);
                        v[i] = div_nelems(t, nelems, inv_nelems);
                    }
#pragma unroll
                    for (int q = 0; q < kD / 4; ++q)
                        *reinterpret_cast<float4 *>(dst + k * kNhwcPixPitch + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                    dst[k * kNhwcPixPitch + kD - 1] = v[kD - 1];
                }
            }
            __syncthreads();
            // one warp per pixel: the group's channels are contiguous in the destination pixel
            for (int pix = warp; pix < kTY * kTX; pix += 8) {
                const int ry = pix / kTX, cx = pix - ry * kTX;
                const int y = 2 * (y0 + ry) + py, x = 2 * (x0 + cx) + px;
                if (y < H && x < W) {
                    const float *src = stage + ry * kNhwcRowPitch + cx * kNhwcPixPitch;
                    float *dst = out + (((size_t)n * H + y) * W + x) * c_dst + c_off + tj0 * kD;
                    for (int ch = lane; ch < n_ch; ch += 32) {
                        const int tjl = ch / kD;                                   // constant divisor: mul + shift
                        const float v = src[ch + tjl * (kNhwcTjPitch - kD)];
                        dst[ch] = v > 0.f ? v : __fmul_rn(v, slope);
                    }
                }
            }
            __syncthreads();
        }
    } else {
#pragma unroll 1
    for (int grp = 0; grp < (kD + kEpiGroup - 1) / kEpiGroup; ++grp) {
        const bool mine = live && tj >= grp * kEpiGroup && tj < (grp + 1) * kEpiGroup;
        const int n_tj = min(kD, (grp + 1) * kEpiGroup) - grp * kEpiGroup;
        if (mine) {
            float *dst = stage + ((tj - grp * kEpiGroup) * kD * kTY + yi) * kEpiPitch + xb * kPX;
#pragma unroll
            for (int i = 0; i < kD; ++i) {
                float v[kPX];
#pragma unroll
                for (int k = 0; k < kPX; ++k) {
                    float t;
                    if (k & 1) t = (i == 0) ? accs[k] : (((i - 1) & 1) ? accp[k][(i - 1) / 2].y : accp[k][(i - 1) / 2].x);
                    else       t = (i == kD - 1) ? accs[k] : ((i & 1) ? accp[k][i / 2].y : accp[k][i / 2].x);
                    v[k] = div_nelems(t, nelems, inv_nelems);
                }
                *reinterpret_cast<float4 *>(dst + i * (kTY * kEpiPitch)) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4 *>(dst + i * (kTY * kEpiPitch) + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
        }
        __syncthreads();
        // each warp takes (tj-in-group, row) pairs and walks the 21 horizontal displacements with
        // constant strides: one LDS + one STG per output row segment
        const int x = 2 * (x0 + lane) + px;
        for (int pr = warp; pr < n_tj * kTY; pr += 8) {
            const int tjl = pr / kTY, ry = pr - tjl * kTY;
            const int y = 2 * (y0 + ry) + py;
            if (y < H && x < W) {
                const float *src = stage + (tjl * (kD * kTY) + ry) * kEpiPitch + lane;
                if constexpr (std::is_same<OT, float>::value) {
                    float *dst = out_n + (size_t)((grp * kEpiGroup + tjl) * kD) * hw + (size_t)y * W + x;
#pragma unroll
                    for (int ti = 0; ti < kD; ++ti) dst[(size_t)ti * hw] = src[ti * (kTY * kEpiPitch)];
                } else {
                    unsigned short *dst = reinterpret_cast<unsigned short *>(out) + (size_t)n * (kD * kD) * hw +
                                          (size_t)((grp * kEpiGroup + tjl) * kD) * hw + (size_t)y * W + x;
#pragma unroll
                    for (int ti = 0; ti < kD; ++ti) dst[(size_t)ti * hw] = Io16<OT>::from_float(src[ti * (kTY * kEpiPitch)]);
                }
            }
        }
        __syncthreads();
    }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int make_plane_map(CUtensorMap *tm, float *base, const CorrGeom &g, int Hp, int pitch, int box_w, int box_h)
{
    const cuuint64_t plane_bytes = (cuuint64_t)Hp * pitch * 4;
    const cuuint64_t dims[4] = {(cuuint64_t)pitch, (cuuint64_t)Hp, (cuuint64_t)g.C, (cuuint64_t)g.B * 4};
    const cuuint64_t strides[3] = {(cuuint64_t)pitch * 4, plane_bytes, plane_bytes * g.C};
    const cuuint32_t box[4] = {(cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)kCK, 1};
    return encode_map4(tm, base, dims, strides, box, "corr_fwd");
}

static int corr_fast_workspace_check(const CorrGeom &g, void *ws, size_t ws_bytes, float *&P1, float *&P2, const char *who)
{
    const PlaneGeom p = plane_geom(g);
    const size_t need = corr_fast_fwd_workspace(g);
    FLOWOPS_REQUIRE(ws && ws_bytes >= need && ((uintptr_t)ws & 255) == 0, FLOWOPS_EWORKSPACE,
                    "%s: workspace of %zu bytes (256-byte aligned) required, got %zu", who, need, ws_bytes);
    P1 = reinterpret_cast<float *>(ws);
    P2 = P1 + (size_t)g.B * 4 * g.C * p.elems1;
    return 0;
}

// NHWC features -> planes of input slot `only` (0: f1, 1: f2; -1: both), optionally fused with the producing
// convolution's bias + LeakyReLU epilogue
int corr_fast_planes_nhwc(const float *in1, const float *in2, const CorrGeom &g, int only, const float *bias, float slope,
                          float *act, void *ws, size_t ws_bytes, cudaStream_t st)
{
    float *P1, *P2;
    int rc = corr_fast_workspace_check(g, ws, ws_bytes, P1, P2, "corr planes");
    if (rc) return rc;
    const PlaneGeom p = plane_geom(g);
    if ((g.W & 7) || (g.H & 1)) {      // planes then contain cells no thread writes
        float *base = only == 1 ? P2 : P1;
        const size_t bytes = only < 0 ? corr_fast_fwd_workspace(g)
                                      : sizeof(float) * (size_t)g.B * 4 * g.C * (only == 1 ? p.elems2 : p.elems1);
        cudaError_t e = cudaMemsetAsync(base, 0, bytes, st);
        if (e != cudaSuccess) { set_error("corr planes: memset failed: %s", cudaGetErrorString(e)); return (int)e; }
    }
    const int c_tiles = (g.C + kNhwcCh - 1) / kNhwcCh, x_tiles_in = (g.W + kNhwcPix - 1) / kNhwcPix;
    const int nz = only < 0 ? g.B * 2 : g.B;
    FLOWOPS_REQUIRE(g.H <= 65535 && nz <= 65535, FLOWOPS_EUNSUPPORTED, "corr planes: NHWC input too large for the transpose grid");
    corr_planarize_nhwc<<<dim3(c_tiles * x_tiles_in, g.H, nz), 256, 0, st>>>(in1, in2, P1, P2, g.C, g.H, g.W, p.Hp, p.Wp,
                                                                           p.pitch1, p.pitch2, only, bias, slope, act);
    return check_launch("corr_planarize_nhwc");
}

// the correlation proper, on planes already in the workspace
int corr_fast_main(float *out, const CorrGeom &g, void *ws, size_t ws_bytes, cudaStream_t st,
                   bool nhwc_out, int c_dst, int c_off, float slope, int out_dtype)
{
    float *P1, *P2;
    int rc = corr_fast_workspace_check(g, ws, ws_bytes, P1, P2, "corr_fwd");
    if (rc) return rc;
    const PlaneGeom p = plane_geom(g);
    CUtensorMap tm1, tm2;
    rc = make_plane_map(&tm1, P1, g, p.Hp, p.pitch1, kF1W, kTY);
    if (rc) return rc;
    rc = make_plane_map(&tm2, P2, g, p.Hp, p.pitch2, kF2W, kF2H);
    if (rc) return rc;

    // cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute: a process that runs the
    // operator on several GPUs must set it on each of them, so it is set before every launch (a host-side call that
    // is cheap, thread-safe and legal during stream capture) for the instantiation about to run.
    const void *fn = nhwc_out ? (const void *)corr_fwd_fast<2, true>
                   : out_dtype == FLOWOPS_DTYPE_F16 ? (const void *)corr_fwd_fast<2, false, __half>
                   : out_dtype == FLOWOPS_DTYPE_BF16 ? (const void *)corr_fwd_fast<2, false, __nv_bfloat16>
                                                     : (const void *)corr_fwd_fast<2, false>;
    {
        const cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) { set_error("corr_fwd: cannot reserve %d bytes of shared memory: %s", kSmemBytes, cudaGetErrorString(e)); return (int)e; }
    }
    const int row_tiles = (p.Hp + kTY - 1) / kTY, x_tiles = (p.Wp + kTX - 1) / kTX;
    const size_t grid = (size_t)g.B * row_tiles * x_tiles * 4;
    FLOWOPS_REQUIRE(grid < (1ull << 31), FLOWOPS_EUNSUPPORTED, "corr_fwd: grid too large");
    if (nhwc_out)
        corr_fwd_fast<2, true><<<(unsigned)grid, 256, kSmemBytes, st>>>(tm1, tm2, out, g.C, g.H, g.W, row_tiles, x_tiles, c_dst, c_off, slope);
    else if (out_dtype == FLOWOPS_DTYPE_F16)
        corr_fwd_fast<2, false, __half><<<(unsigned)grid, 256, kSmemBytes, st>>>(tm1, tm2, out, g.C, g.H, g.W, row_tiles, x_tiles, 0, 0, 1.f);
    else if (out_dtype == FLOWOPS_DTYPE_BF16)
        corr_fwd_fast<2, false, __nv_bfloat16><<<(unsigned)grid, 256, kSmemBytes, st>>>(tm1, tm2, out, g.C, g.H, g.W, row_tiles, x_tiles, 0, 0, 1.f);
    else
        corr_fwd_fast<2, false><<<(unsigned)grid, 256, kSmemBytes, st>>>(tm1, tm2, out, g.C, g.H, g.W, row_tiles, x_tiles, 0, 0, 1.f);
    return check_launch("corr_fwd_fast");
}

// io_dtype 0: fp32 tensors (flowops_corr_fwd); FLOWOPS_DTYPE_F16 / _BF16: 16-bit inputs and output, NCHW only
// (flowops_corr_fwd_16) -- in1, in2 and out then point at 16-bit elements
int corr_fast_fwd_launch(const float *in1, const float *in2, float *out, const CorrGeom &g, int in_layout,
                         void *ws, size_t ws_bytes, cudaStream_t st, int io_dtype)
{
    float *P1, *P2;
    int rc = corr_fast_workspace_check(g, ws, ws_bytes, P1, P2, "corr_fwd");
    if (rc) return rc;
    if (in_layout == FLOWOPS_LAYOUT_NHWC) {
        FLOWOPS_REQUIRE(io_dtype == 0, FLOWOPS_EUNSUPPORTED, "corr_fwd: channels-last inputs are fp32 only");
        rc = corr_fast_planes_nhwc(in1, in2, g, -1, nullptr, 0.f, nullptr, ws, ws_bytes, st);
        if (rc) return rc;
    } else {
        const PlaneGeom p = plane_geom(g);
        if ((g.W & 7) || (g.H & 1)) {      // planes have zero padding only when W is not a multiple of 8 or H is odd
            cudaError_t e = cudaMemsetAsync(ws, 0, corr_fast_fwd_workspace(g), st);
            if (e != cudaSuccess) { set_error("corr_fwd: memset failed: %s", cudaGetErrorString(e)); return (int)e; }
        }
        // fp32 rows of W % 4 == 0 floats from a 16-byte-aligned base are 16-byte aligned; 16-bit rows are 8-byte aligned
        const int vec_ok = (g.W % 4 == 0) && (io_dtype == 0 ? (aligned16(in1) && aligned16(in2))
                                                            : ((((uintptr_t)in1 | (uintptr_t)in2) & 7u) == 0));
        const size_t total = (size_t)g.B * g.C * g.H * ((g.W + 3) / 4);
        size_t blocks = (total + 255) / 256;
        if (blocks > (size_t)kNumSMs * 8 * 8) blocks = (size_t)kNumSMs * 8 * 8;
        const dim3 pgrid((unsigned)blocks, 2);
        if (io_dtype == FLOWOPS_DTYPE_F16)
            corr_planarize<__half><<<pgrid, 256, 0, st>>>(reinterpret_cast<const __half *>(in1), reinterpret_cast<const __half *>(in2),
                                                          P1, P2, g.B, g.C, g.H, g.W, p.Hp, p.pitch1, p.pitch2, vec_ok);
        else if (io_dtype == FLOWOPS_DTYPE_BF16)
            corr_planarize<__nv_bfloat16><<<pgrid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16 *>(in1), reinterpret_cast<const __nv_bfloat16 *>(in2),
                                                                 P1, P2, g.B, g.C, g.H, g.W, p.Hp, p.pitch1, p.pitch2, vec_ok);
        else
            corr_planarize<float><<<pgrid, 256, 0, st>>>(in1, in2, P1, P2, g.B, g.C, g.H, g.W, p.Hp, p.pitch1, p.pitch2, vec_ok);
        rc = check_launch("corr_planarize");
        if (rc) return rc;
    }
    return corr_fast_main(out, g, ws, ws_bytes, st, false, 0, 0, 1.f, io_dtype);
}

}  // namespace flowops
