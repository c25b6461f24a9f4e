// corr_tc_bwd.cu -- FlowNetC Correlation backward on the Blackwell tensor cores (tcgen05 / TMEM / TMA), sm_100a.
//
//   gI1[n,c,y,x] = (1/C) sum_{tj,ti} gO[n,(tj,ti),y,x]                      * f2pad[n,c, y+2(tj-10), x+2(ti-10)]
//   gI2[n,c,y,x] = (1/C) sum_{tj,ti} gO[n,(tj,ti),y-2(tj-10),x-2(ti-10)]    * f1[n,c,y-2(tj-10),x-2(ti-10)]
//   (reference correlation_cuda_kernel.cu:151-334; terms whose gO coordinate is out of range are dropped)
//
// Both are the banded contraction  out[p][c] = sum_d G[d][p] * Win[p + d][c]  inside one parity plane (corr_bwd.cu):
//   gI1: G = gO, Win = f2;      gI2: G[d'][q] = gO[-d'][q + d'], Win = f1.
//
// The contraction as a GEMM.  For a tile of 16 x 8 = 128 pixels and the (16+20) x (8+20) = 36 x 28 window of Win positions
//       D[m, c] = sum_w  A[m, w] * Bw[w, c]        A[m = (r, cx), w = (wy, wx)] = G[(wy - r, wx - cx)][pixel m] inside the band, else 0
//                                                  Bw = the window's features, N = C channels (<= 256 per pass)
// i.e. M = 128, N = C, K = 1008 window positions of which 441 per row are non-zero (43.75 % of the dense MMA is useful, as
// in the forward kernel, corr_tc.cu).  The accumulator (128 lanes x C fp32 columns) stays in TMEM for the whole tile; TMEM
// holds two, so the epilogue of one tile runs under the UMMAs of the next.
//
// Operands.  Bw comes by TMA from "P32" planes P[n*4+parity][c/32][Y][X][c%32] written by a layout pass: a K block is 2
// window rows x 4 window columns; the box (32 channels, 4 X, 2 Y, C/32) lands in shared memory as [c/32][8 positions][32
// channels] -- 128-byte rows of channels, the MN-major SWIZZLE_128B_BASE32B UMMA layout that tf32 operands require.  A is
// built by the "skew" warps from pixel-major records of G that a pre-pass writes: G'[pixel][tj][wx_local] with the row of
// 21 horizontal displacements already shifted to window coordinates (28 entries, zeros outside the band), so that a thread
// owning (pixel m, window row parity) fetches its 4 K-block entries with one aligned 128-bit load and only the vertical band
// limit is a predicate.  3xTF32 as in the forward: hi = the fp32 word (the tensor core truncates), lo = x - trunc(x) rounded
// to TF32, D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi with fp32 accumulation.
//
// Roles (one persistent CTA per SM, 448 threads): warp 0 = TMA producer (Bw), warp 1 = TMEM allocation + UMMA issue (one
// thread), warps 2-9 = skew + split (A hi / lo tiles from G', lo tile of Bw), warps 10-13 = epilogue (tcgen05.ld -> 1/C ->
// NCHW store).  mbarriers: full (TMA -> split), ready (skew / split -> UMMA), empty (tcgen05.commit -> TMA and skew),
// tmem_full[2] / tmem_empty[2] (UMMA <-> epilogue); every wait is bounded.
#include <stdlib.h>

#include "corr.cuh"
#include "tc.cuh"

namespace flowops {
namespace tcb {

constexpr int kD = 21, kR = 10;
constexpr int M = 128;                                      // pixels per tile = UMMA M
constexpr int KB_Y = 2, KB_X = 4;                           // a K block: 2 window rows x 4 window columns = K of one tf32 UMMA
// Pixel tile (plane rows x plane columns) and its window of (TH + 20) x (TW + 20) positions = 126 K blocks either way:
// 16 x 8 -> 36 x 28 (18 x 7 K blocks), 8 x 16 -> 28 x 36 (14 x 9).  The host picks the shape with fewer (partial) tiles.
template <int TH_, int TW_> struct Tile {
    static constexpr int TH = TH_, TW = TW_;
    static constexpr int WH = TH + 2 * kR, WW = TW + 2 * kR;
    static constexpr int NA = WH / KB_Y, NB = WW / KB_X;
    static_assert(TH * TW == M && WH % KB_Y == 0 && WW % KB_X == 0 && KB_Y * KB_X == 8, "K blocks tile the window");
};
constexpr int A_BYTES = M * 8 * 4;                          // 4096
constexpr int B_MAX = 256 * 8 * 4;                          // 8192 (N = 256 channels)
constexpr int N_SPLIT_WARPS = 4, N_SKEW_WARPS = 8, N_EPI_WARPS = 4;
constexpr int THREADS = 32 * (2 + N_SPLIT_WARPS + N_SKEW_WARPS + N_EPI_WARPS);      // 576
constexpr int SMEM_BARRIERS = 512;

using tc::mbar_wait_b;

// MN-major shared-memory matrix descriptor for 32-bit operands.  tf32 operands with the MN dimension contiguous have ONE
// legal layout, SWIZZLE_128B_BASE32B (cute::UMMA::Layout_MN_SW128_32B_Atom; with b_major = MN and any other layout type the
// UMMA silently produces zeros -- measured): an atom is 4 K rows of 128 bytes (32 channels), 32-byte chunks XOR-swizzled with
// the row index (Swizzle<2,5,2> on the byte address).  Stride byte offset = distance between the two K atoms of a K = 8 UMMA
// (512 B: the second window row of the K block), leading byte offset = distance between 32-channel groups (1024 B).
__device__ __forceinline__ uint64_t smem_desc_mn_sw128_32b(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(1024 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (1ull << 61);
}

struct Params {
    const float *gout;      // gO [B, 441, H, W]
    float *out[2];          // gI1, gI2 (NCHW); a null entry is never selected (which0 / n_which)
    int which0, n_which;
    int C, NC, n_chunks;    // channels, channels per pass (UMMA N), passes
    int nacc, acc_stride;   // accumulators in TMEM (2 when two fit next to the A ring) and their column stride
    int H, W, PH, PW;
    int tilesY, tilesX, planes, n_items;
    float nelems, inv_nelems;
    int flags;              // bit 1 (as in the forward): single TF32 product (layout debugging)
    unsigned long long *trace;   // debugging: per-CTA wait counters (8 per CTA), or null
};

struct Item { int which, plane, Y0, X0, chunk; };
template <class T>
__device__ __forceinline__ Item decode_item(int item, const Params &p)
{
    Item it;
    it.chunk = item % p.n_chunks; item /= p.n_chunks;
    const int tx = item % p.tilesX; item /= p.tilesX;
    const int ty = item % p.tilesY; item /= p.tilesY;
    it.plane = item % p.planes;
    it.which = p.which0 + item / p.planes;
    it.Y0 = ty * T::TH; it.X0 = tx * T::TW;
    return it;
}

// TMEM loads / stores and the A-from-TMEM UMMA used only here
__device__ __forceinline__ void tmem_st4(uint32_t taddr, float4 v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 :: "r"(taddr), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]: A is 128 lanes x 8 columns of tf32 words
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}"
        :: "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ float4 tf32_residual(float4 v)
{
    // lo = x - trunc_tf32(x) (exact), rounded to TF32 by adding half an ulp to the bit pattern (the tensor core truncates)
    float4 l;
    l.x = __uint_as_float(__float_as_uint(__fsub_rn(v.x, __uint_as_float(__float_as_uint(v.x) & 0xffffe000u))) + 0x1000u);
    l.y = __uint_as_float(__float_as_uint(__fsub_rn(v.y, __uint_as_float(__float_as_uint(v.y) & 0xffffe000u))) + 0x1000u);
    l.z = __uint_as_float(__float_as_uint(__fsub_rn(v.z, __uint_as_float(__float_as_uint(v.z) & 0xffffe000u))) + 0x1000u);
    l.w = __uint_as_float(__float_as_uint(__fsub_rn(v.w, __uint_as_float(__float_as_uint(v.w) & 0xffffe000u))) + 0x1000u);
    return l;
}

// ATM = true: the A operand (skewed gradient tile, hi and lo) lives in TMEM (tcgen05.st by the skew warps, 16 columns per
// stage next to ONE 256-column accumulator) and shared memory carries only Bw; ATM = false: A tiles in shared memory
// (K-major SWIZZLE_32B) and two accumulators.
template <bool ATM> struct Cfg {
    static constexpr int STAGES = ATM ? 12 : 8;
    static constexpr int B_OFS = ATM ? 0 : 2 * A_BYTES;
    static constexpr int STAGE_BYTES = B_OFS + 2 * B_MAX;       // [A hi, A lo,] B raw (= hi), B lo
    static constexpr int A_COL0 = 256;                          // first TMEM column of the A ring (ATM)
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + SMEM_BARRIERS + 1024;
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    static_assert(!ATM || A_COL0 + 16 * STAGES <= 512, "TMEM budget");
};

template <bool ATM, class T>
__global__ void __launch_bounds__(THREADS, 1)
corr_bwd_tc(const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1, const Params p)
{
    using K = Cfg<ATM>;
    constexpr int TW = T::TW, WW = T::WW, NA = T::NA, NB = T::NB;
    constexpr int STAGES = K::STAGES, STAGE_BYTES = K::STAGE_BYTES, B_OFS = K::B_OFS;
    const int NACC = p.nacc;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar_base = base + STAGES * STAGE_BYTES;
    auto bar_full = [&](int s) { return bar_base + 8u * s; };
    auto bar_ready = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto bar_empty = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
    auto bar_tfull = [&](int a) { return bar_base + 8u * (3 * STAGES + a); };
    auto bar_tempty = [&](int a) { return bar_base + 8u * (3 * STAGES + 2 + a); };
    const uint32_t slot_addr = bar_base + 8u * (3 * STAGES + 4);
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen + (slot_addr - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b_bytes = 32 * p.NC;                                   // one K block of Bw

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_ready(s), 1 + N_SKEW_WARPS);                // one split warp per K block (round robin) + the skew warps
            mbar_init(bar_empty(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull(a), 1);
            mbar_init(bar_tempty(a), N_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(slot_addr), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer: one K block of the window's features per stage =================
        if (lane == 0) {
            uint32_t it = 0;
            long long w_empty = 0;
            const long long t_start = clock64();
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                const Item w = decode_item<T>(item, p);
                const CUtensorMap *tm = w.which == 0 ? &tmB0 : &tmB1;
                for (int a = 0; a < NA; ++a)
                    for (int b = 0; b < NB; ++b, ++it) {
                        const int s = it % STAGES;
                        const uint32_t u = it / STAGES;
                        w_empty += mbar_wait_b(bar_empty(s), (u & 1) ^ 1);
                        mbar_expect_tx(bar_full(s), b_bytes);
                        // dims (c%32, X, Y, c/32, plane): lands as [c/32][wy_local][wx_local][32 channels]; zero fill outside the plane
                        tc::tma_load_5d(base + s * STAGE_BYTES + B_OFS, tm, 0, w.X0 - kR + KB_X * b, w.Y0 - kR + KB_Y * a,
                                        w.chunk * (p.NC >> 5), w.plane, bar_full(s));
                    }
            }
            if (p.trace) { p.trace[blockIdx.x * 8 + 0] = w_empty; p.trace[blockIdx.x * 8 + 6] = clock64() - t_start; }
        }
    } else if (warp == 1) {
        // ================= UMMA issuer =================
        if (lane == 0) {
            // D = fp32, A and B = TF32, A K-major, B MN-major (bit 16), N >> 3 in bits 17-22, M >> 4 in bits 24-28
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(p.NC >> 3) << 17) | ((128u >> 4) << 24);
            uint32_t it = 0, j = 0;
            long long w_ready = 0, w_tempty = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++j) {
                const int acc = j % NACC;
                w_tempty += mbar_wait_b(bar_tempty(acc), ((j / NACC) & 1) ^ 1);
                tc::tc_fence_after();
                const uint32_t d = tmem + (uint32_t)(p.acc_stride * acc);
                for (int kb = 0; kb < NA * NB; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t u = it / STAGES;
                    w_ready += mbar_wait_b(bar_ready(s), u & 1);
                    tc::tc_fence_after();
                    const uint32_t sb = base + s * STAGE_BYTES;
                    const uint64_t b_hi = smem_desc_mn_sw128_32b(sb + B_OFS), b_lo = smem_desc_mn_sw128_32b(sb + B_OFS + B_MAX);
                    const uint32_t first = kb == 0 ? 0u : 1u;
                    if (ATM) {
                        const uint32_t a_hi = tmem + K::A_COL0 + 16u * s, a_lo = a_hi + 8u;
                        if (!(p.flags & 2)) {
                            umma_tf32_ts(d, a_lo, b_hi, idesc, first);
                            umma_tf32_ts(d, a_hi, b_lo, idesc, 1u);
                            umma_tf32_ts(d, a_hi, b_hi, idesc, 1u);
                        } else {
                            umma_tf32_ts(d, a_hi, b_hi, idesc, first);       // single TF32 product: timing / layout debugging only
                        }
                    } else {
                        const uint64_t a_hi = tc::smem_desc_sw32(sb), a_lo = tc::smem_desc_sw32(sb + A_BYTES);
                        tc::umma_tf32(d, a_lo, b_hi, idesc, first);
                        tc::umma_tf32(d, a_hi, b_lo, idesc, 1u);
                        tc::umma_tf32(d, a_hi, b_hi, idesc, 1u);
                    }
                    tc::tc_commit(bar_empty(s));
                }
                tc::tc_commit(bar_tfull(acc));
            }
            if (p.trace) { p.trace[blockIdx.x * 8 + 1] = w_ready; p.trace[blockIdx.x * 8 + 2] = w_tempty; }
        }
    } else if (warp < 2 + N_SPLIT_WARPS) {
        // ================= split: lo tile of Bw, same position as the raw tile; split warp k takes K blocks k, k + 4, ... =================
        const int sw = warp - 2;
        const int n_chunks16 = 2 * p.NC;                             // 16-byte chunks of one Bw K block
        const uint32_t n_kb = (uint32_t)((p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * (NA * NB);
        long long w_full = 0, t_split = 0;
        for (uint32_t it = sw; it < n_kb; it += N_SPLIT_WARPS) {
            const int s = it % STAGES;
            const uint32_t u = it / STAGES;
            uint8_t *braw = gen + s * STAGE_BYTES + B_OFS;
            w_full += mbar_wait_b(bar_full(s), u & 1);               // Bw has landed (and the stage's previous readers are done)
            const long long ts0 = clock64();
#pragma unroll 1
            for (int i0 = 0; i0 < n_chunks16; i0 += 128) {
                float4 v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = *reinterpret_cast<const float4 *>(braw + (i0 + lane + 32 * k) * 16);
#pragma unroll
                for (int k = 0; k < 4; ++k) *reinterpret_cast<float4 *>(braw + B_MAX + (i0 + lane + 32 * k) * 16) = tf32_residual(v[k]);
            }
            tc::fence_proxy_async();                                 // generic-proxy writes -> visible to the UMMA (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_ready(s));
            t_split += clock64() - ts0;
        }
        if (p.trace && sw == 0 && lane == 0) { p.trace[blockIdx.x * 8 + 3] = w_full; p.trace[blockIdx.x * 8 + 5] = t_split; }
    } else if (warp < 2 + N_SPLIT_WARPS + N_SKEW_WARPS) {
        // ================= skew: the banded gradient tile A (hi and lo) from gO =================
        const int m = 32 * (warp & 3) + lane;                        // accumulator row = TMEM lane (a warp reaches its own lane quarter)
        const int yy = (warp - (2 + N_SPLIT_WARPS)) >> 2;            // window row parity inside the K block
        const int r = m / TW, cx = m % TW;                           // m = r * TW + cx
        const uint32_t a_ofs = (uint32_t)m * 32u + 16u * (uint32_t)(yy ^ ((m >> 2) & 1));     // 32-byte swizzle: chunk ^= address bit 7
        const uint32_t a_lane = (uint32_t)(32 * (warp & 3)) << 16;
        const size_t hw = (size_t)p.H * p.W;
        uint32_t it = 0;
        long long w_sk_empty = 0, t_sk = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            const Item w = decode_item<T>(item, p);
            const int Y = w.Y0 + r, X = w.X0 + cx;
            const bool pix_ok = Y < p.PH && X < p.PW;
            const int n = w.plane >> 2, py = (w.plane >> 1) & 1, px = w.plane & 1;
            const float *go_n = p.gout + (size_t)n * (kD * kD) * hw;
            // A[m][(wy, wx)] = G[(tj = wy - r, ti = wx - cx)][my pixel]: per window row 21 gradient values, one per horizontal
            // displacement, placed at window columns cx .. cx + 20 of the row's 28.  The loads are issued per displacement (all
            // lanes read the same channel plane: 8 pixels of a tile row share two sectors) and the lane-dependent shift by cx
            // is applied in registers when the row is consumed (one select level per bit of cx).
            auto load_row = [&](int a, float (&v)[kD]) {
                const int tj = KB_Y * a + yy - r;
#pragma unroll
                for (int ti = 0; ti < kD; ++ti) v[ti] = 0.f;
                if (!(pix_ok && tj >= 0 && tj < kD)) return;
                if (w.which == 0) {
                    // gI1: G = gO: channel tj*21 + ti at my pixel; consecutive ti are one channel plane apart
                    const float *src = go_n + (size_t)(tj * kD) * hw + (size_t)(2 * Y + py) * p.W + (2 * X + px);
#pragma unroll
                    for (int ti = 0; ti < kD; ++ti) v[ti] = ldg_stream(src + (size_t)ti * hw);
                } else {
                    // gI2: G[(tj, ti)][q] = gO[(20 - tj, 20 - ti)][q + (tj - 10, ti - 10)], zero outside the plane
                    const int Ys = Y + tj - kR;
                    if (Ys < 0 || Ys >= p.PH) return;
                    const float *src = go_n + (size_t)((kD - 1 - tj) * kD + kD - 1) * hw + (size_t)(2 * Ys + py) * p.W + px + 2 * (X - kR);
                    const ptrdiff_t step = 2 - (ptrdiff_t)hw;            // ti + 1: one channel down, one plane column right
#pragma unroll
                    for (int ti = 0; ti < kD; ++ti) {
                        const int Xs = X + ti - kR;
                        if (Xs >= 0 && Xs < p.PW) v[ti] = ldg_stream(src + (ptrdiff_t)ti * step);
                    }
                }
            };
            float vcur[kD], vnext[kD];
            load_row(0, vcur);
#pragma unroll 1
            for (int a = 0; a < NA; ++a) {
                if (a + 1 < NA) load_row(a + 1, vnext);              // the next window row's values are in flight while this one is consumed
                float wv[WW];                                        // the row in window coordinates: wv[j] = v[j - cx]
#pragma unroll
                for (int j = 0; j < WW; ++j) wv[j] = j < kD ? vcur[j] : 0.f;
#pragma unroll
                for (int bit = 1; bit < TW; bit <<= 1) {
                    const bool on = (cx & bit) != 0;
#pragma unroll
                    for (int j = WW - 1; j >= 0; --j) wv[j] = on ? (j >= bit ? wv[j - bit] : 0.f) : wv[j];
                }
#pragma unroll
                for (int b = 0; b < NB; ++b, ++it) {
                    const int s = it % STAGES;
                    const uint32_t u = it / STAGES;
                    w_sk_empty += mbar_wait_b(bar_empty(s), (u & 1) ^ 1);          // the UMMAs that read this stage's A tiles are done
                    const long long ts0 = clock64();
                    const float4 v = make_float4(wv[4 * b], wv[4 * b + 1], wv[4 * b + 2], wv[4 * b + 3]), l = tf32_residual(v);
                    if (ATM) {
                        tc::tc_fence_after();
                        const uint32_t col = tmem + a_lane + K::A_COL0 + 16u * s + 4u * yy;
                        tmem_st4(col, v);
                        tmem_st4(col + 8u, l);
                        tmem_st_wait();
                        tc::tc_fence_before();
                    } else {
                        uint8_t *stage = gen + s * STAGE_BYTES;
                        *reinterpret_cast<float4 *>(stage + a_ofs) = v;
                        *reinterpret_cast<float4 *>(stage + A_BYTES + a_ofs) = l;
                        tc::fence_proxy_async();
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_ready(s));
                    t_sk += clock64() - ts0;
                }
#pragma unroll
                for (int ti = 0; ti < kD; ++ti) vcur[ti] = vnext[ti];
            }
        }
        if (p.trace && warp == 2 + N_SPLIT_WARPS && lane == 0) { p.trace[blockIdx.x * 8 + 4] = w_sk_empty; p.trace[blockIdx.x * 8 + 7] = t_sk; }
    } else {
        // ================= epilogue: TMEM -> 1/C -> NCHW =================
        const int q = warp & 3;                                      // TMEM lane quarter this warp may read
        const int m = 32 * q + lane;
        const int r = m / TW, cx = m % TW;
        const size_t hw = (size_t)p.H * p.W;
        uint32_t j = 0;
        long long t_epi = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++j) {
            const Item w = decode_item<T>(item, p);
            const int acc = j % NACC;
            const int n = w.plane >> 2, py = (w.plane >> 1) & 1, px = w.plane & 1;
            const int Y = w.Y0 + r, X = w.X0 + cx;
            const bool pix_ok = Y < p.PH && X < p.PW;
            float *dst = p.out[w.which] + ((size_t)n * p.C + (size_t)w.chunk * p.NC) * hw + (size_t)(2 * Y + py) * p.W + (2 * X + px);
            mbar_wait_b(bar_tfull(acc), (j / NACC) & 1);
            const long long te0 = clock64();
            tc::tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < p.NC; c0 += 16) {
                float v[16];
                tc::tmem_ld16(tmem + (uint32_t)(p.acc_stride * acc) + ((uint32_t)(32 * q) << 16) + (uint32_t)c0, v);
                tc::tmem_ld_wait();
                if (pix_ok) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) stg_stream(dst + (size_t)(c0 + i) * hw, div_nelems(v[i], p.nelems, p.inv_nelems));
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty(acc));
            t_epi += clock64() - te0;
        }
        (void)t_epi;
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------
// layout pass
// ---------------------------------------------------------------------------------------------------------------
// NCHW -> "P32" planes P[n*4 + parity][c/32][Y][X][c%32]: 128-byte rows of 32 channels per plane position, the row format of
// the MN-major tf32 operand (a 128-byte inner TMA dimension is what CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B takes; with the
// forward's 32-byte P8 rows that swizzle mode faults -- measured, tools/tma_sw_probe.cu).  A thread owns one pixel and one
// group of 32 channels: 32 coalesced row reads (lane = x), one 128-byte row written.
__global__ void __launch_bounds__(256) planes32_from_nchw(const float *__restrict__ in1, const float *__restrict__ in2,
                                                          float *__restrict__ P1, float *__restrict__ P2, int C, int H, int W)
{
    const int CG = C >> 5, PH = H >> 1, PW = W >> 1;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int z = blockIdx.z;                                    // (n, input, cg)
    const int cg = z % CG, second = (z / CG) & 1, n = (z / CG) >> 1;
    if (x >= W || y >= H) return;
    const size_t hw = (size_t)H * W;
    const float *in = (second ? in2 : in1) + ((size_t)n * C + cg * 32) * hw + (size_t)y * W + x;
    float v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = ldg_stream(in + (size_t)k * hw);
    const int par = (y & 1) * 2 + (x & 1);
    float *dst = (second ? P2 : P1) + (((((size_t)n * 4 + par) * CG + cg) * PH + (y >> 1)) * PW + (x >> 1)) * 32;
#pragma unroll
    for (int k = 0; k < 8; ++k) stg_stream4(dst + 4 * k, make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
}

}  // namespace tcb

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static inline size_t up256(size_t n) { return (n + 255) & ~(size_t)255; }
static size_t tcb_plane_bytes(const CorrGeom &g) { return up256(sizeof(float) * (size_t)g.B * g.C * g.H * g.W); }

bool corr_tc_bwd_supported(const CorrGeom &g)
{
    return corr_tc_supported(g) && (g.C <= 256 || g.C % 256 == 0) && g.H <= 65535 && g.B <= 65535;
}

size_t corr_tc_bwd_workspace(const CorrGeom &g)
{
    return 2 * tcb_plane_bytes(g);        // P32 planes of both inputs
}

int corr_tc_bwd_launch(const float *in1, const float *in2, const float *gout, float *gin1, float *gin2,
                       const CorrGeom &g, void *ws, size_t ws_bytes, cudaStream_t st)
{
    const size_t need = corr_tc_bwd_workspace(g);
    FLOWOPS_REQUIRE(ws && ws_bytes >= need && ((uintptr_t)ws & 255) == 0, FLOWOPS_EWORKSPACE,
                    "corr_bwd: workspace of %zu bytes (256-byte aligned) required, got %zu", need, ws_bytes);
    float *P1 = reinterpret_cast<float *>(ws);
    float *P2 = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(ws) + tcb_plane_bytes(g));
    {   // P32 planes of both inputs
        const dim3 grid((g.W + 31) / 32, (g.H + 7) / 8, 2 * g.B * (g.C / 32));
        FLOWOPS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, FLOWOPS_EUNSUPPORTED, "corr_bwd: input too large for the layout grid");
        tcb::planes32_from_nchw<<<grid, 256, 0, st>>>(in1, in2, P1, P2, g.C, g.H, g.W);
        const int rc0 = check_launch("planes32_from_nchw");
        if (rc0) return rc0;
    }
    int rc = 0;

    const int PH = g.H / 2, PW = g.W / 2;
    // Channels per pass (UMMA N).  Passes of 128 channels would halve the work items (a finer tail) and let two accumulators
    // sit next to the A ring in TMEM, but a tf32 UMMA of N = 128 takes as long as one of N = 256 (measured: 178 vs 190 clocks),
    // so the full width is the default; FLOWOPS_TCB_NC overrides it for A/B timing.
    int NC = g.C <= 256 ? g.C : 256;
    if (const char *e = getenv("FLOWOPS_TCB_NC")) { const int v = atoi(e); if (v > 0 && v <= 256 && v % 32 == 0 && g.C % v == 0) NC = v; }   // A/B timing
    CUtensorMap tmB0, tmB1;
    const cuuint64_t row = 128, line = (cuuint64_t)PW * row, img = line * PH, plane = img * (g.C / 32);
    const cuuint64_t dims[5] = {32, (cuuint64_t)PW, (cuuint64_t)PH, (cuuint64_t)(g.C / 32), (cuuint64_t)g.B * 4};
    const cuuint64_t strides[4] = {row, line, img, plane};
    const cuuint32_t box[5] = {32, tcb::KB_X, tcb::KB_Y, (cuuint32_t)(NC / 32), 1};
    rc = encode_map5_sw32(&tmB0, P2, dims, strides, box, "corr_bwd", CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);     // gI1 contracts gO with f2
    if (rc) return rc;
    rc = encode_map5_sw32(&tmB1, P1, dims, strides, box, "corr_bwd", CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);     // gI2 contracts the mirrored gO with f1
    if (rc) return rc;

    tcb::Params p;
    p.gout = gout;
    p.out[0] = gin1; p.out[1] = gin2;
    p.which0 = gin1 ? 0 : 1;
    p.n_which = (gin1 ? 1 : 0) + (gin2 ? 1 : 0);
    p.C = g.C; p.NC = NC; p.n_chunks = g.C / NC;
    const bool atm = !(((corr_impl_flags() >> 1) & 3) & 1);
    p.nacc = (!atm || NC <= 128) ? 2 : 1;
    p.acc_stride = (atm && NC <= 128) ? 128 : 256;
    p.H = g.H; p.W = g.W; p.PH = PH; p.PW = PW;
    // tile shape: the one that needs fewer (partial) tiles, 16 x 8 on a tie (e.g. 24 x 32 planes: 6 tiles of 8 x 16, not 8 of 16 x 8)
    const long long t168 = (long long)((PH + 15) / 16) * ((PW + 7) / 8), t816 = (long long)((PH + 7) / 8) * ((PW + 15) / 16);
    const bool wide = t816 < t168;
    p.tilesY = wide ? (PH + 7) / 8 : (PH + 15) / 16;
    p.tilesX = wide ? (PW + 15) / 16 : (PW + 7) / 8;
    p.planes = g.B * 4;
    p.n_items = p.n_which * p.planes * p.tilesY * p.tilesX * p.n_chunks;
    p.nelems = (float)g.C; p.inv_nelems = 1.f / (float)g.C;
    p.flags = (corr_impl_flags() >> 1) & 3;
    p.trace = corr_tc_trace_buffer();
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = p.n_items < sms ? p.n_items : sms;
    // p.flags & 1 (flowops_corr_set_impl(3)): A tiles in shared memory (the first version; kept for A/B timing)
    using T168 = tcb::Tile<16, 8>;
    using T816 = tcb::Tile<8, 16>;
    void (*kern)(const CUtensorMap, const CUtensorMap, const tcb::Params) =
        atm ? (wide ? tcb::corr_bwd_tc<true, T816> : tcb::corr_bwd_tc<true, T168>)
            : (wide ? tcb::corr_bwd_tc<false, T816> : tcb::corr_bwd_tc<false, T168>);
    const int smem = atm ? tcb::Cfg<true>::SMEM_BYTES : tcb::Cfg<false>::SMEM_BYTES;
    // per-device attribute, set before every launch (see corr_fast.cu)
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("corr_bwd: cannot reserve %d bytes of shared memory: %s", smem, cudaGetErrorString(e)); return (int)e; }
    kern<<<grid, tcb::THREADS, smem, st>>>(tmB0, tmB1, p);
    return check_launch("corr_bwd_tc");
}

}  // namespace flowops
