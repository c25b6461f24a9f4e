// flowhead.cu -- the decoders' 2-channel flow heads as a direct FP32 convolution.
//
// predict_flow (reference networks/submodules.py:40-41: nn.Conv2d(in_planes, 2, kernel_size=3, stride=1, padding=1))
// runs 16 times per FlowNet2 forward, on 16 .. 1026 input channels.  Two output channels are a poor fit for an implicit
// GEMM: cuDNN pads the filter count in front of every call, un-pads the result behind it (nhwcAddPaddingKernel: 84 us per
// 16 pairs for the full-resolution head of the fusion network) and leaves the bias to a third pass -- 382 us for a layer
// that moves 604 MB (94 us at the HBM copy rate).  Here the layer is what it looks like: per output pixel 2 x 9 x Cin
// FP32 multiply-adds on channels-last activations, bias included, no temporary.
//
// Work split.  A warp owns a strip 32 pixels wide and `rows` rows tall and streams down it: lane = (pixel group g = lane / 4,
// channel quad q = lane % 4); a thread handles its four channels of every 16-channel chunk for the four pixels 4g .. 4g+3.
//   * Every INPUT row is loaded once per strip -- the four lanes of a pixel group read the 64 contiguous bytes of a pixel's
//     chunk with one 128-bit load each (16 fully used sectors per request) -- and feeds the three output rows it touches,
//     whose partial sums live in registers (3 rows x 4 pixels x 2 outputs); a finished row is summed over the quads with two
//     xor-shuffles and stored, and the accumulators rotate.
//   * The loads of step s + 1 (next chunk, or the next row's first chunk) are issued before the multiply-adds of step s
//     into the other of two register buffers: there is no barrier in the loop, so a warp always has a row segment in flight.  (A first version that synchronised
//     the CTA per chunk and had no prefetch ran at 415 us on the full-resolution head -- latency-bound at 16 warps per SM.)
//   * All weights sit in shared memory for the whole CTA (72 x 16 bytes per chunk) and are read as 128-bit words shared by
//     the four pixels of a thread; the arithmetic is packed FFMA2 on (out0, out1) pairs.
#include <stdlib.h>

#include "common.cuh"
#include "tma.cuh"

namespace flowops {

constexpr int kHeadWarps = 8;          // strips per CTA (stacked vertically)
constexpr int kHeadChunkQuads = 72;    // float4s of weights per 16-channel chunk: 9 taps x 4 quads x 2

struct HeadRowLoad {
    float4 v[6];
};

// One input row's worth of multiply-adds of one 16-channel chunk: acc[k] is output row (r - 1 + k), which sees input row r
// through tap row dy = 2 - k.  W(tap, half) yields the thread's 8 weights of a tap as two float4s.
template <typename WF>
__device__ __forceinline__ void head_row_fma(const HeadRowLoad &cur, float2 (&acc)[3][4], WF W)
{
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const float4 wa = W((2 - k) * 3 + dx, 0), wb = W((2 - k) * 3 + dx, 1);
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const float4 t = cur.v[p + dx];
                acc[k][p] = fma2(make_float2(t.x, t.x), make_float2(wa.x, wa.y), acc[k][p]);
                acc[k][p] = fma2(make_float2(t.y, t.y), make_float2(wa.z, wa.w), acc[k][p]);
                acc[k][p] = fma2(make_float2(t.z, t.z), make_float2(wb.x, wb.y), acc[k][p]);
                acc[k][p] = fma2(make_float2(t.w, t.w), make_float2(wb.z, wb.w), acc[k][p]);
            }
        }
    }
}

// Input row r is consumed: output row r - 1 has received its last contribution -- and, if r is the image's last row, so has
// row r.  Sum over the four channel quads, add the bias, lane q stores pixel q of its group; then the accumulators rotate.
__device__ __forceinline__ void head_finish_row(float2 (&acc)[3][4], int r, int H, int y0, int y1, int x0, int q, int W,
                                                float b0, float b1, float2 *ob)
{
#pragma unroll
    for (int fin = 0; fin < 2; ++fin) {
        const int yo = r - 1 + fin;
        if (fin == 1 && r != H - 1) break;
        float2 mine = make_float2(0.f, 0.f);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            float sx = acc[fin][p].x, sy = acc[fin][p].y;
            sx += __shfl_xor_sync(0xffffffffu, sx, 1); sy += __shfl_xor_sync(0xffffffffu, sy, 1);
            sx += __shfl_xor_sync(0xffffffffu, sx, 2); sy += __shfl_xor_sync(0xffffffffu, sy, 2);
            if (p == q) mine = make_float2(sx + b0, sy + b1);
        }
        if (yo >= y0 && yo < y1 && x0 + q < W) ob[(size_t)yo * W + x0 + q] = mine;
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) { acc[0][p] = acc[1][p]; acc[1][p] = acc[2][p]; acc[2][p] = make_float2(0.f, 0.f); }
}

// wp: [n_chunks][tap = dy*3+dx][quad][channel in quad][out] (zero beyond the layer's real channels)
// 120 registers, two CTAs per SM (a build aimed at three spills and runs at half the speed)
__global__ void __launch_bounds__(32 * kHeadWarps, 2) flow_head_kernel(const float *__restrict__ x, const float *__restrict__ wp,
                                                                       const float *__restrict__ bias, float2 *__restrict__ out,
                                                                       int H, int W, unsigned c_pitch, int cin, int n_chunks, int rows)
{
    extern __shared__ float4 sw[];
    for (int i = threadIdx.x; i < n_chunks * kHeadChunkQuads; i += blockDim.x) sw[i] = __ldg(reinterpret_cast<const float4 *>(wp) + i);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int x0 = blockIdx.x * 32 + g * 4;
    const int y0 = (blockIdx.y * kHeadWarps + warp) * rows;
    if (y0 >= H) return;                                   // whole warps only; no barrier below
    const int y1 = min(y0 + rows, H);
    const size_t b = blockIdx.z;
    const float *xb = x + b * (size_t)H * W * c_pitch + q * 4;
    float2 *ob = out + b * (size_t)H * W;
    const float b0 = bias ? __ldg(bias) : 0.f, b1 = bias ? __ldg(bias + 1) : 0.f;
    bool col_ok[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) col_ok[j] = x0 - 1 + j >= 0 && x0 - 1 + j < W;
    const bool quad_tail = n_chunks * 16 - 16 + q * 4 >= cin;      // this quad of the LAST chunk lies beyond the input channels

    // input rows y0-1 .. y1 (the out-of-image ones contribute nothing and are skipped), n_chunks steps per row
    const int r_first = max(y0 - 1, 0), r_last = min(y1, H - 1);
    auto load = [&](int r, int chunk, HeadRowLoad &t) {
        const bool ok = !(quad_tail && chunk == n_chunks - 1);
        const long long base = ((long long)r * W + (x0 - 1)) * (long long)c_pitch + chunk * 16;       // x0 - 1 may be -1 (never read)
#pragma unroll
        for (int j = 0; j < 6; ++j)
            t.v[j] = (ok && col_ok[j]) ? __ldg(reinterpret_cast<const float4 *>(xb + (base + (long long)j * c_pitch))) : make_float4(0.f, 0.f, 0.f, 0.f);
    };

    // acc[k][p]: output row (r - 1 + k) while input row r is being consumed: k = 0 gets tap row dy = 2, k = 1 dy = 1, k = 2 dy = 0
    float2 acc[3][4];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[k][p] = make_float2(0.f, 0.f);

    // steps (r, chunk) in order; two register buffers used alternately, so the loads of step s + 1 are in flight while step s
    // is multiplied and nothing waits for them before step s + 1 starts (a `cur = nxt` copy would)
    HeadRowLoad bufA, bufB;
    int r = r_first, chunk = 0;
    load(r, 0, bufA);
    auto step = [&](const HeadRowLoad &cur, HeadRowLoad &nxt) -> bool {
        const bool last_chunk = chunk + 1 == n_chunks;
        const int rn = last_chunk ? r + 1 : r, cn = last_chunk ? 0 : chunk + 1;
        const bool more = rn <= r_last;
        if (more) load(rn, cn, nxt);
        const float4 *w = sw + chunk * kHeadChunkQuads + q * 2;
        head_row_fma(cur, acc, [&](int tap, int half) { return w[tap * 8 + half]; });
        if (last_chunk) head_finish_row(acc, r, H, y0, y1, x0, q, W, b0, b1, ob);
        r = rn;
        chunk = cn;
        return more;
    };
    while (step(bufA, bufB) && step(bufB, bufA)) {
    }
}

}  // namespace flowops

extern "C" int flowops_flow_head_nhwc(const float *x, int c_pitch, int cin, const float *w_packed, const float *bias, float *out,
                                      int B, int H, int W, void *stream)
{
    using namespace flowops;
    FLOWOPS_REQUIRE(x && w_packed && out, FLOWOPS_EINVAL, "flow_head_nhwc: null pointer");
    FLOWOPS_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && cin > 0 && (cin & 3) == 0 && c_pitch >= cin && (c_pitch & 3) == 0,
                    FLOWOPS_EINVAL, "flow_head_nhwc: bad shape (%d x %d x %d, %d channels at a pitch of %d; multiples of 4)", B, H, W, cin, c_pitch);
    FLOWOPS_REQUIRE(aligned16(x) && aligned16(w_packed) && (reinterpret_cast<uintptr_t>(out) & 7u) == 0, FLOWOPS_EINVAL,
                    "flow_head_nhwc: x and w_packed must be 16-byte aligned, out 8-byte aligned");
    const int n_chunks = (cin + 15) / 16;
    const size_t smem = (size_t)n_chunks * kHeadChunkQuads * sizeof(float4);
    constexpr size_t kHeadMaxSmem = 200 * 1024;
    FLOWOPS_REQUIRE(smem <= kHeadMaxSmem, FLOWOPS_EUNSUPPORTED, "flow_head_nhwc: %d input channels need %zu bytes of shared memory", cin, smem);
    // rows per strip: tall strips re-read fewer halo rows, short ones give the small decoder levels enough warps (1.5 x the
    // 148 SMs x 16 resident ones)
    const long strips_x = (W + 31) / 32;
    int rows = 32;
    while (rows > 2 && strips_x * ((H + rows - 1) / rows) * B < 3L * kNumSMs * kHeadWarps) rows >>= 1;
    if (const char *t = getenv("FLOWOPS_TUNE_HEAD_ROWS")) {        // A/B timing only
        const int v = atoi(t);
        if (v >= 1 && v <= 512) rows = v;
    }
    const dim3 grid((unsigned)strips_x, (unsigned)((H + rows * kHeadWarps - 1) / (rows * kHeadWarps)), (unsigned)B);
    FLOWOPS_REQUIRE(grid.y <= 65535, FLOWOPS_EINVAL, "flow_head_nhwc: H %d too large", H);
    // per device, cheap, legal during stream capture: set on every call rather than caching a per-process flag -- and always
    // to the same ceiling, not to this call's size, so that two host threads launching heads of different widths cannot
    // lower the limit under each other between this call and the launch
    cudaError_t e = cudaFuncSetAttribute(flow_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHeadMaxSmem);
    FLOWOPS_REQUIRE(e == cudaSuccess, (int)e, "flow_head_nhwc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    flow_head_kernel<<<grid, 32 * kHeadWarps, smem, (cudaStream_t)stream>>>(x, w_packed, bias, reinterpret_cast<float2 *>(out), H, W,
                                                                            (unsigned)c_pitch, cin, n_chunks, rows);
    return check_launch("flow_head_nhwc");
}
