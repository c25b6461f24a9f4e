// corr_tc.cu -- FlowNetC Correlation forward on the Blackwell tensor cores (tcgen05 / TMEM / TMA), sm_100a.
//
//   out[n, tj*21+ti, y, x] = (1/C) * sum_c f1[n,c,y,x] * f2pad[n,c, y + 2(tj-10), x + 2(ti-10)]
//   (reference correlation_cuda_kernel.cu:74-147; FlowNetC.py:31: pad 20, k 1, md 20, s1 1, s2 2)
//
// The contraction as a GEMM.  stride2 = 2 splits the problem into four parity planes in which the displacement is a
// dense +-10 x +-10 window (see corr_fast.cu).  Inside one plane, for a tile of 16 x 8 = 128 pixels,
//       D[m, n] = sum_c  A[m, c] * B[n, c]        A = the tile's f1 pixels (M = 128 rows),
//                                                  B = the (16+20) x (8+20) = 36 x 28 window of f2 positions,
// holds every output of the tile: pixel (r, c) needs the 21 x 21 window positions (r..r+20, c..c+20), i.e. 441 of the
// 1008 columns of its row of D (43.75 % of the dense tile is useful -- the price of running a band-structured
// contraction on a dense MMA; the FLOP counts reported for this kernel are the useful ones, 2*441*C per pixel).
// A work item is (pixel tile, quarter of the window): N = 9 rows x 28 = 252 columns = one UMMA N = 256 (4 padding
// columns), so TMEM (128 lanes x 512 fp32 columns) holds two accumulators and the epilogue of one item overlaps the
// UMMAs of the next.
//
// Precision: 3xTF32.  x = hi + lo with hi = x truncated to TF32 (what the tensor core reads from an fp32 word) and
// lo = x - hi (exact), D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi, fp32 accumulation in TMEM: max-relative error ~1e-6
// against the fp32 reference kernel (tolerance 1e-5; tools/split_mma_probe.py, tests/test_corr_tc_gpu.py).  The raw
// fp32 tile brought in by TMA *is* the hi operand; "splitter" warps derive the lo tile from it in shared memory, so
// L2 -> SM traffic is one fp32 copy of the data.
//
// Data layout.  Operands must be K-major (channels contiguous) for the tf32 UMMA; the producer of the features writes
// "P8" planes  P[n*4 + parity][c/8][Y][X][c%8]  (32-byte rows = the K = 8 of one tf32 UMMA; consecutive X contiguous,
// so a TMA box row is one 32-byte sector and a box line is a contiguous run).  Zero padding of the window comes from
// TMA's out-of-bounds fill.  A box lands in shared memory as consecutive 32-byte rows with the 32-byte swizzle, which is
// the canonical K-major SWIZZLE_32B UMMA layout (8-row core matrices, SBO = 256 B).
//
// Roles (one persistent CTA per SM, 576 threads): warp 0 = TMA producer, warp 1 = TMEM allocation + UMMA issue (one
// thread), warps 2-9 = splitters, warps 10-17 = epilogue (tcgen05.ld -> scale / LeakyReLU -> shared-memory transpose ->
// channels-last store: each pixel's run of channels is contiguous; for the reference's NCHW layout the accumulator rows are
// ordered tile-row-major instead and stored straight from registers, 8- or 16-pixel row segments per channel).  mbarrier pipelines: full (TMA -> splitters),
// split (splitters -> UMMA), empty (tcgen05.commit -> TMA), tmem_full[2] / tmem_empty[2] (UMMA <-> epilogue).
#include "corr.cuh"
#include "tma.cuh"
#include "tc.cuh"

namespace flowops {
namespace tc {

constexpr int kD = 21, kR = 10;
constexpr int M = 128;                          // pixels per tile = UMMA M
constexpr int NUSED = 252;                      // accumulator columns in use: a quarter of the window (both tile shapes)
constexpr int NPAD = 256;                       // one UMMA N = 256
constexpr int KC = 8;                           // channels per pipeline stage = K of one tf32 UMMA (32 bytes)
constexpr int STAGES = 6;
constexpr int A_BYTES = M * KC * 4;             // 4096
constexpr int B_BYTES = NPAD * KC * 4;          // 8192 (the TMA box fills NUSED rows = 8064 bytes)
constexpr int B_TX = NUSED * KC * 4;
constexpr int RAW_BYTES = A_BYTES + B_BYTES;    // 12288: raw (= hi) tile of A, then of B
constexpr int STAGE_BYTES = 2 * RAW_BYTES;      // raw + lo
constexpr int SPLIT_CHUNKS = (A_BYTES + B_TX) / 16;   // 760 16-byte chunks to split per stage
constexpr int N_SPLIT_WARPS = 8, N_EPI_WARPS = 8;
constexpr int SPLIT_PER_THREAD = (SPLIT_CHUNKS + 32 * N_SPLIT_WARPS - 1) / (32 * N_SPLIT_WARPS);   // 3
constexpr int THREADS = 32 * (2 + N_SPLIT_WARPS + N_EPI_WARPS);     // 576
constexpr int EPI_STG_BYTES = 32 * kD * 4;      // one window row of a warp's 32 pixels: 32 x 21 floats
constexpr int SMEM_BARRIERS = 256;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + N_EPI_WARPS * EPI_STG_BYTES + SMEM_BARRIERS + 1024;   // + alignment slack
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

// Pixel tile (plane rows x plane columns, TH * TW = 128) and its window: (TH + 20) x (TW + 20) f2 positions, processed in
// four quarters of QROWS window rows.  16 x 8: window 36 x 28, quarters of 9 rows; 8 x 16: window 28 x 36, quarters of 7
// rows -- 252 accumulator columns either way.  The host picks the shape that wastes fewer partial tiles.
template <int TH_, int TW_> struct Tile {
    static constexpr int TH = TH_, TW = TW_;
    static constexpr int WH = TH + 2 * kR, WW = TW + 2 * kR;
    static constexpr int QROWS = WH / 4;
    static_assert(TH * TW == M && WH % 4 == 0 && QROWS * WW == NUSED, "tile shape");
};

struct Params {
    float *out;
    int CB;                 // channel blocks of 8
    int H, W, PH, PW;       // image and plane extents (H, W even)
    int tilesY, tilesX, n_items;
    int c_dst, c_off;       // channels-last destination: [B, H, W, c_dst], channels [c_off, c_off + 441)
    float slope, nelems, inv_nelems;
    unsigned long long *trace;   // debugging: per-CTA wait / work cycle counters (8 per CTA), or null
    int flags;              // bit 0: splitters also rewrite the hi tile with its low 13 mantissa bits cleared
                            // bit 1: single TF32 product (layout debugging; not within tolerance)
};

struct Item { int plane, Y0, X0, h; };          // h: which quarter of the window (rows 9h .. 9h+8)
template <class T>
__device__ __forceinline__ Item decode_item(int item, const Params &p)
{
    Item it;
    it.h = item & 3;
    int t = item >> 2;
    const int tx = t % p.tilesX; t /= p.tilesX;
    const int ty = t % p.tilesY;
    it.plane = t / p.tilesY;
    it.Y0 = ty * T::TH; it.X0 = tx * T::TW;
    return it;
}

// T: tile shape.  NCHW = false: channels-last store (LeakyReLU folded in) through a shared-memory transpose, A rows arrive
// as m = x*TH + y.  NCHW = true: the reference's [B, 441, H, W] layout stored straight from registers, A rows arrive as
// m = y*TW + x so that a warp holds whole tile rows (8 or 16 pixels of one output row per store segment).
template <class T, bool NCHW>
__global__ void __launch_bounds__(THREADS, 1)
corr_fwd_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p)
{
    constexpr int QROWS = T::QROWS, WW = T::WW, TW = T::TW;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));           // generic-address view of the aligned base
    const uint32_t stg_base = base + STAGES * STAGE_BYTES;
    const uint32_t bar_base = stg_base + N_EPI_WARPS * EPI_STG_BYTES;
    // barriers: full[s], split[s], empty[s], tmem_full[2], tmem_empty[2]; then the TMEM base address
    auto bar_full = [&](int s) { return bar_base + 8u * s; };
    auto bar_split = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto bar_empty = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
    auto bar_tfull = [&](int a) { return bar_base + 8u * (3 * STAGES + a); };
    auto bar_tempty = [&](int a) { return bar_base + 8u * (3 * STAGES + 2 + a); };
    const uint32_t slot_addr = bar_base + 8u * (3 * STAGES + 4);
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen + (slot_addr - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_split(s), N_SPLIT_WARPS);
            mbar_init(bar_empty(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull(a), 1);
            mbar_init(bar_tempty(a), N_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(slot_addr), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= 2 && warp < 2 + N_SPLIT_WARPS) {
        // the 4 padding rows of every B tile (raw and lo) are read by the UMMA and never written by TMA: keep them
        // finite so that the (unused) padding columns of the accumulator cannot hold NaNs
        const int t = threadIdx.x - 64;
        constexpr int per = (B_BYTES - B_TX) / 4;                    // floats per padding block
        for (int i = t; i < STAGES * 2 * per; i += 32 * N_SPLIT_WARPS) {
            const int blk = i / per, o = i - blk * per;
            const int s = blk >> 1, which = blk & 1;
            *reinterpret_cast<float *>(gen + s * STAGE_BYTES + which * RAW_BYTES + A_BYTES + B_TX + o * 4) = 0.f;
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t it = 0;
            long long w_empty = 0;
            const long long t_start = clock64();
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                const Item w = decode_item<T>(item, p);
                for (int kb = 0; kb < p.CB; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t u = it / STAGES;
                    w_empty += mbar_wait_b(bar_empty(s), (u & 1) ^ 1);
                    mbar_expect_tx(bar_full(s), A_BYTES + B_TX);
                    const uint32_t dst = base + s * STAGE_BYTES;
                    // A: map dims (c%8, Y, X, ...) -> rows m = x_local * TH + y_local; NCHW: (c%8, X, Y, ...) -> m = y_local * TW + x_local
                    if (NCHW) tma_load_5d(dst, &tmA, 0, w.X0, w.Y0, kb, w.plane, bar_full(s));
                    else      tma_load_5d(dst, &tmA, 0, w.Y0, w.X0, kb, w.plane, bar_full(s));
                    // B: dims (c%8, X, Y, c/8, plane): rows n = wy_local * 28 + wx_local; zero fill outside the plane
                    tma_load_5d(dst + A_BYTES, &tmB, 0, w.X0 - kR, w.Y0 - kR + QROWS * w.h, kb, w.plane, bar_full(s));
                }
            }
            if (p.trace) { p.trace[blockIdx.x * 8 + 0] = w_empty; p.trace[blockIdx.x * 8 + 6] = clock64() - t_start; }
        }
    } else if (warp == 1) {
        // ================= UMMA issuer =================
        if (lane == 0) {
            uint32_t it = 0, j = 0;
            long long w_split = 0, w_tempty = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++j) {
                const int acc = j & 1;                               // accumulator buffer (TMEM columns 256 * acc ...)
                w_tempty += mbar_wait_b(bar_tempty(acc), ((j >> 1) & 1) ^ 1);    // the epilogue has drained its previous use
                tc_fence_after();
                const uint32_t d = tmem + 256u * acc;
                for (int kb = 0; kb < p.CB; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t u = it / STAGES;
                    w_split += mbar_wait_b(bar_split(s), u & 1);
                    tc_fence_after();
                    const uint32_t sb = base + s * STAGE_BYTES;
                    const uint64_t a_hi = smem_desc_sw32(sb), b_hi = smem_desc_sw32(sb + A_BYTES);
                    const uint64_t a_lo = smem_desc_sw32(sb + RAW_BYTES), b_lo = smem_desc_sw32(sb + RAW_BYTES + A_BYTES);
                    const uint32_t first = kb == 0 ? 0u : 1u;
                    if (!(p.flags & 2)) {
                        umma_tf32(d, a_lo, b_hi, kIdesc, first);
                        umma_tf32(d, a_hi, b_lo, kIdesc, 1u);
                        umma_tf32(d, a_hi, b_hi, kIdesc, 1u);
                    } else {
                        umma_tf32(d, a_hi, b_hi, kIdesc, first);
                    }
                    tc_commit(bar_empty(s));                         // stage reusable when these UMMAs have read it
                }
                tc_commit(bar_tfull(acc));                           // accumulator complete
            }
            if (p.trace) { p.trace[blockIdx.x * 8 + 1] = w_split; p.trace[blockIdx.x * 8 + 2] = w_tempty; }
        }
    } else if (warp < 2 + N_SPLIT_WARPS) {
        // ================= splitters: lo = x - trunc_tf32(x), same position in the lo tile =================
        const int t = threadIdx.x - 64;
        uint32_t it = 0;
        long long w_full = 0, t_work = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            for (int kb = 0; kb < p.CB; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t u = it / STAGES;
                w_full += mbar_wait_b(bar_full(s), u & 1);
                const long long tw0 = clock64();
                uint8_t *raw = gen + s * STAGE_BYTES;
                float4 v[SPLIT_PER_THREAD];
#pragma unroll
                for (int k = 0; k < SPLIT_PER_THREAD; ++k) {         // all loads first: one shared-memory latency per stage
                    const int i = t + k * 32 * N_SPLIT_WARPS;
                    if (i < SPLIT_CHUNKS) v[k] = *reinterpret_cast<const float4 *>(raw + i * 16);
                }
#pragma unroll
                for (int k = 0; k < SPLIT_PER_THREAD; ++k) {
                    const int i = t + k * 32 * N_SPLIT_WARPS;
                    if (i < SPLIT_CHUNKS) {
                        float4 h, l;
                        h.x = __uint_as_float(__float_as_uint(v[k].x) & 0xffffe000u);
                        h.y = __uint_as_float(__float_as_uint(v[k].y) & 0xffffe000u);
                        h.z = __uint_as_float(__float_as_uint(v[k].z) & 0xffffe000u);
                        h.w = __uint_as_float(__float_as_uint(v[k].w) & 0xffffe000u);
                        // the residual has at most 13 significant bits; the tensor core keeps 11 of them by truncation:
                        // round to nearest first (+ half an ulp of TF32 on the bit pattern)
                        l.x = __uint_as_float(__float_as_uint(__fsub_rn(v[k].x, h.x)) + 0x1000u);
                        l.y = __uint_as_float(__float_as_uint(__fsub_rn(v[k].y, h.y)) + 0x1000u);
                        l.z = __uint_as_float(__float_as_uint(__fsub_rn(v[k].z, h.z)) + 0x1000u);
                        l.w = __uint_as_float(__float_as_uint(__fsub_rn(v[k].w, h.w)) + 0x1000u);
                        *reinterpret_cast<float4 *>(raw + RAW_BYTES + i * 16) = l;
                        if (p.flags & 1) *reinterpret_cast<float4 *>(raw + i * 16) = h;
                    }
                }
                fence_proxy_async();                                 // generic-proxy writes -> visible to the UMMA (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_split(s));
                t_work += clock64() - tw0;
            }
        }
        if (p.trace && t == 0) { p.trace[blockIdx.x * 8 + 3] = w_full; p.trace[blockIdx.x * 8 + 7] = t_work; }
    } else {
        // ================= epilogue =================
        const int ew = warp - (2 + N_SPLIT_WARPS);                   // 0..7
        const int q = warp & 3;                                      // TMEM lane quarter this warp may read
        const int eh = ew >> 2;                                      // even / odd window rows of the item
        const int m = 32 * q + lane;                                 // accumulator row
        uint32_t j = 0;
        long long w_tfull = 0, t_epi = 0;
        if constexpr (!NCHW) {
            const int r = m % T::TH, c = m / T::TH;                  // m = x_local * TH + y_local
            constexpr int LPC = T::TH;                               // lanes per tile column inside a warp
            const int sel = lane / LPC;                              // column of this lane relative to the warp's first one
            constexpr int NSEL = 32 / LPC;                           // 2 (16 x 8 tile) or 4 (8 x 16)
            constexpr int NLOAD = kD + NSEL - 1 <= 24 ? 24 : 32;
            float *stg = reinterpret_cast<float *>(gen + (stg_base - base) + ew * EPI_STG_BYTES);
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++j) {
                const Item w = decode_item<T>(item, p);
                const int acc = j & 1;
                const int n = w.plane >> 2, py = (w.plane >> 1) & 1, px = w.plane & 1;
                const int Y = w.Y0 + r, X = w.X0 + c;
                const bool pix_ok = Y < p.PH && X < p.PW;
                const int pix_ofs = ((n * p.H + 2 * Y + py) * p.W + 2 * X + px) * p.c_dst + p.c_off;
                w_tfull += mbar_wait_b(bar_tfull(acc), (j >> 1) & 1);
                const long long te0 = clock64();
                tc_fence_after();
#pragma unroll 1
                for (int wl = eh; wl < QROWS; wl += 2) {             // window row of this item handled now
                    const int tj = QROWS * w.h + wl - r;             // vertical displacement index of that row for my pixel
                    float v[NLOAD];
                    const uint32_t taddr = tmem + 256u * acc + ((uint32_t)(32 * q) << 16) + (uint32_t)(wl * WW + NSEL * q);
                    tmem_ld16(taddr, v);
                    if (NLOAD == 24) tmem_ld8(taddr + 16, v + 16); else tmem_ld16(taddr + 16, v + 16);
                    tmem_ld_wait();
                    __syncwarp();                                    // previous row's staging fully read
#pragma unroll
                    for (int i = 0; i < kD; ++i) {
                        float t = v[i];
                        if (NSEL == 2) t = sel ? v[i + 1] : v[i];
                        else t = sel == 0 ? v[i] : (sel == 1 ? v[i + 1] : (sel == 2 ? v[i + 2] : v[i + 3]));
                        t = div_nelems(t, p.nelems, p.inv_nelems);
                        t = t > 0.f ? t : __fmul_rn(t, p.slope);
                        stg[lane * kD + i] = t;
                    }
                    const int my_ofs = (pix_ok && tj >= 0 && tj < kD) ? pix_ofs + tj * kD : -1;
                    __syncwarp();
#pragma unroll
                    for (int e = 0; e < kD; ++e) {                   // 32 x 21 values, 32 per trip, pixel-major
                        const int f = e * 32 + lane;
                        const int pl = f / kD, i = f - pl * kD;
                        const int ofs = __shfl_sync(0xffffffffu, my_ofs, pl);
                        if (ofs >= 0) p.out[ofs + i] = stg[f];
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty(acc));
                t_epi += clock64() - te0;
            }
        } else {
            // NCHW: m = y_local * TW + x_local, so this warp holds 32 / TW whole tile rows; for a fixed channel a store
            // instruction writes 32 / TW row segments of TW pixels (stride 2 in the image: the other column parity's tile
            // fills the gaps).  The per-lane column offset c is taken out with a select tree (no shared-memory staging).
            const int r = m / TW, c = m % TW;
            constexpr int NLOAD = (kD + TW - 1 + 3) / 4 * 4;         // 28 (TW = 8) or 36 (TW = 16) columns of the window row
            const size_t hw = (size_t)p.H * p.W;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++j) {
                const Item w = decode_item<T>(item, p);
                const int acc = j & 1;
                const int n = w.plane >> 2, py = (w.plane >> 1) & 1, px = w.plane & 1;
                const int Y = w.Y0 + r, X = w.X0 + c;
                const bool pix_ok = Y < p.PH && X < p.PW;
                float *pix = p.out + (size_t)n * (kD * kD) * hw + (size_t)(2 * Y + py) * p.W + (2 * X + px);
                w_tfull += mbar_wait_b(bar_tfull(acc), (j >> 1) & 1);
                const long long te0 = clock64();
                tc_fence_after();
#pragma unroll 1
                for (int wl = eh; wl < QROWS; wl += 2) {
                    const int tj = QROWS * w.h + wl - r;
                    float v[40];
                    const uint32_t taddr = tmem + 256u * acc + ((uint32_t)(32 * q) << 16) + (uint32_t)(wl * WW);
                    tmem_ld16(taddr, v);
                    if (NLOAD == 28) { tmem_ld8(taddr + 16, v + 16); tmem_ld8(taddr + 20, v + 20); }      // columns 0..27 (20..23 twice)
                    else { tmem_ld16(taddr + 16, v + 16); tmem_ld8(taddr + 28, v + 28); }                // columns 0..35
                    tmem_ld_wait();
                    // v[i] <- v[i + c]: one select level per bit of c
#pragma unroll
                    for (int bit = 1; bit < TW; bit <<= 1) {
                        const bool on = (c & bit) != 0;
#pragma unroll
                        for (int i = 0; i < NLOAD - bit; ++i) v[i] = on ? v[i + bit] : v[i];
                    }
                    if (pix_ok && tj >= 0 && tj < kD) {
                        float *dst = pix + (size_t)(tj * kD) * hw;
#pragma unroll
                        for (int i = 0; i < kD; ++i) {
                            stg_stream(dst, div_nelems(v[i], p.nelems, p.inv_nelems));
                            dst += hw;
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty(acc));
                t_epi += clock64() - te0;
            }
        }
        if (p.trace && ew == 0 && lane == 0) { p.trace[blockIdx.x * 8 + 4] = w_tfull; p.trace[blockIdx.x * 8 + 5] = t_epi; }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------
// plane writers: P[n*4 + py*2 + px][c/8][Y][X][c%8]
// ---------------------------------------------------------------------------------------------------------------
// NCHW -> P8.  A thread owns one pixel and one block of 8 channels: 8 coalesced row reads (lane = x), one 32-byte write;
// even and odd lanes write two contiguous 512-byte runs (the two column parities).
__global__ void __launch_bounds__(256) planes8_from_nchw(const float *__restrict__ in1, const float *__restrict__ in2,
                                                         float *__restrict__ P1, float *__restrict__ P2,
                                                         int C, int H, int W)
{
    const int CB = C >> 3, PH = H >> 1, PW = W >> 1;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int z = blockIdx.z;                                    // (input, n, cb)
    const int cb = z % CB, n = (z / CB) >> 1, second = (z / CB) & 1;
    if (x >= W || y >= H) return;
    const float *in = (second ? in2 : in1) + (((size_t)n * C + cb * 8) * H + y) * W + x;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = ldg_stream(in + (size_t)k * H * W);
    const int par = (y & 1) * 2 + (x & 1);
    float *dst = (second ? P2 : P1) + (((((size_t)n * 4 + par) * CB + cb) * PH + (y >> 1)) * PW + (x >> 1)) * 8;
    stg_stream4(dst, make_float4(v[0], v[1], v[2], v[3]));
    stg_stream4(dst + 4, make_float4(v[4], v[5], v[6], v[7]));
}

// NHWC -> P8, optionally as the bias + LeakyReLU epilogue of the convolution that produced the features (and writing
// the activated features back in NHWC order for another consumer).  A CTA transposes 64 pixels of one image row x 32
// channels through shared memory: 128-byte coalesced reads per pixel, 1 KB contiguous writes per (parity, channel block).
// `in` / `act` may alias (in-place epilogue): no __restrict__, plain loads.
constexpr int kPix = 64, kCh = 32, kTilePitch = 36;
__global__ void __launch_bounds__(256) planes8_from_nhwc(const float *in, float *__restrict__ P, int C, int H, int W,
                                                         const float *__restrict__ bias, float slope, float *act)
{
    __shared__ __align__(16) float tile[2][kPix / 2][kTilePitch];   // [column parity][plane column][channel]
    const int c_tiles = C / kCh;
    const int ct = blockIdx.x % c_tiles, xt = blockIdx.x / c_tiles;
    const int y = blockIdx.y, n = blockIdx.z, c0 = ct * kCh, x0 = xt * kPix;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t row_off = ((size_t)n * H + y) * W * C;
    const float b = bias ? __ldg(bias + c0 + lane) : 0.f;
#pragma unroll
    for (int i = 0; i < kPix / 8; ++i) {
        const int pxl = warp + 8 * i, x = x0 + pxl;
        float v = 0.f;
        if (x < W) {
            v = in[row_off + (size_t)x * C + c0 + lane];
            if (bias) {
                const float t = __fadd_rn(v, b);
                v = t > 0.f ? t : __fmul_rn(t, slope);
                if (act) act[row_off + (size_t)x * C + c0 + lane] = v;
            }
        }
        tile[pxl & 1][pxl >> 1][lane] = v;
    }
    __syncthreads();
    // warp = (column parity, channel block of 8): lane = plane column, one 32-byte row each
    const int par_x = warp >> 2, cbl = warp & 3;
    const int CB = C >> 3, PH = H >> 1, PW = W >> 1;
    const int X = (x0 >> 1) + lane;
    if (X < PW) {
        const float4 lo = *reinterpret_cast<const float4 *>(&tile[par_x][lane][cbl * 8]);
        const float4 hi = *reinterpret_cast<const float4 *>(&tile[par_x][lane][cbl * 8 + 4]);
        float *dst = P + (((((size_t)n * 4 + (y & 1) * 2 + par_x) * CB + (c0 >> 3) + cbl) * PH + (y >> 1)) * PW + X) * 8;
        stg_stream4(dst, lo);
        stg_stream4(dst + 4, hi);
    }
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static unsigned long long *g_tc_trace = nullptr;      // debugging aid, see flowops_corr_tc_trace
static int g_corr_impl = -1;      // -1: not read yet; bit 0: tensor-core path on; bits 1-2 -> kernel flags (debug)

unsigned long long *corr_tc_trace_buffer() { return g_tc_trace; }

int corr_impl_flags()
{
    if (g_corr_impl < 0) {
        const char *e = getenv("FLOWOPS_CORR_IMPL");
        g_corr_impl = (e && (e[0] == 'f' || e[0] == '0')) ? 0 : 1;       // "ffma" / "0" selects the FP32-FMA kernel
    }
    return g_corr_impl;
}

bool corr_tc_supported(const CorrGeom &g)
{
    return (corr_impl_flags() & 1) && corr_fast_supported(g) && (g.C % 32) == 0 && (g.H % 2) == 0 && (g.W % 2) == 0 &&
           (size_t)g.B * g.H * g.W * 512 < (1ull << 31);
}

static size_t tc_plane_bytes(const CorrGeom &g) { return sizeof(float) * (size_t)g.B * g.C * g.H * g.W; }

size_t corr_tc_fwd_workspace(const CorrGeom &g, bool /*nchw_out*/)
{
    return 2 * tc_plane_bytes(g);          // the two P8 plane sets; both output layouts are stored by the kernel itself
}

static int tc_ws(const CorrGeom &g, void *ws, size_t ws_bytes, bool nchw_out, float *&P1, float *&P2, float *&tmp, const char *who)
{
    const size_t need = corr_tc_fwd_workspace(g, nchw_out);
    FLOWOPS_REQUIRE(ws && ws_bytes >= need && ((uintptr_t)ws & 255) == 0, FLOWOPS_EWORKSPACE,
                    "%s: workspace of %zu bytes (256-byte aligned) required, got %zu", who, need, ws_bytes);
    P1 = reinterpret_cast<float *>(ws);
    P2 = P1 + tc_plane_bytes(g) / 4;
    tmp = P2 + tc_plane_bytes(g) / 4;
    return 0;
}

int corr_tc_planes_nchw(const float *in1, const float *in2, const CorrGeom &g, void *ws, size_t ws_bytes, cudaStream_t st)
{
    float *P1, *P2, *tmp;
    int rc = tc_ws(g, ws, ws_bytes, false, P1, P2, tmp, "corr planes");
    if (rc) return rc;
    const dim3 grid((g.W + 31) / 32, (g.H + 7) / 8, 2 * g.B * (g.C / 8));
    FLOWOPS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, FLOWOPS_EUNSUPPORTED, "corr planes: input too large for the layout grid");
    tc::planes8_from_nchw<<<grid, 256, 0, st>>>(in1, in2, P1, P2, g.C, g.H, g.W);
    return check_launch("planes8_from_nchw");
}

int corr_tc_planes_nhwc(const float *in, int which, const CorrGeom &g, const float *bias, float slope, float *act,
                        void *ws, size_t ws_bytes, cudaStream_t st)
{
    float *P1, *P2, *tmp;
    int rc = tc_ws(g, ws, ws_bytes, false, P1, P2, tmp, "corr planes");
    if (rc) return rc;
    const dim3 grid((g.C / tc::kCh) * ((g.W + tc::kPix - 1) / tc::kPix), g.H, g.B);
    FLOWOPS_REQUIRE(g.H <= 65535 && g.B <= 65535, FLOWOPS_EUNSUPPORTED, "corr planes: NHWC input too large for the transpose grid");
    tc::planes8_from_nhwc<<<grid, 256, 0, st>>>(in, which ? P2 : P1, g.C, g.H, g.W, bias, slope, act);
    return check_launch("planes8_from_nhwc");
}

namespace tc {

// the correlation proper on P8 planes in the workspace; out is channels-last [B, H, W, c_dst] (channels c_off..c_off+440,
// LeakyReLU(slope) folded in) or, with NCHW, the reference's [B, 441, H, W]
template <class T, bool NCHW>
static int tc_launch(float *out, const CorrGeom &g, float *P1, float *P2, cudaStream_t st, int c_dst, int c_off, float slope)
{
    const int PH = g.H / 2, PW = g.W / 2, CB = g.C / 8;
    CUtensorMap tmA, tmB;
    const cuuint64_t row = 32, line = (cuuint64_t)PW * row, img = line * PH, plane = img * CB;
    const cuuint64_t dimsYX[5] = {8, (cuuint64_t)PH, (cuuint64_t)PW, (cuuint64_t)CB, (cuuint64_t)g.B * 4};      // Y before X: rows m = x*TH + y
    const cuuint64_t strYX[4] = {line, row, img, plane};
    const cuuint64_t dimsXY[5] = {8, (cuuint64_t)PW, (cuuint64_t)PH, (cuuint64_t)CB, (cuuint64_t)g.B * 4};      // natural order: rows m = y*W_box + x
    const cuuint64_t strXY[4] = {row, line, img, plane};
    const cuuint32_t boxA_yx[5] = {8, T::TH, T::TW, 1, 1}, boxA_xy[5] = {8, T::TW, T::TH, 1, 1};
    const cuuint32_t boxB[5] = {8, T::WW, T::QROWS, 1, 1};
    int rc = NCHW ? encode_map5_sw32(&tmA, P1, dimsXY, strXY, boxA_xy, "corr_fwd") : encode_map5_sw32(&tmA, P1, dimsYX, strYX, boxA_yx, "corr_fwd");
    if (rc) return rc;
    rc = encode_map5_sw32(&tmB, P2, dimsXY, strXY, boxB, "corr_fwd");
    if (rc) return rc;
    Params p;
    p.out = out;
    p.CB = CB; p.H = g.H; p.W = g.W; p.PH = PH; p.PW = PW;
    p.tilesY = (PH + T::TH - 1) / T::TH; p.tilesX = (PW + T::TW - 1) / T::TW;
    p.n_items = g.B * 4 * p.tilesY * p.tilesX * 4;
    p.c_dst = c_dst; p.c_off = c_off; p.slope = slope;
    p.nelems = (float)g.C; p.inv_nelems = 1.f / (float)g.C;
    p.flags = (corr_impl_flags() >> 1) & 3;
    p.trace = g_tc_trace;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // per-device attribute, set before every launch (see corr_fast.cu)
    const cudaError_t e = cudaFuncSetAttribute(corr_fwd_tc<T, NCHW>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) { set_error("corr_fwd: cannot reserve %d bytes of shared memory: %s", SMEM_BYTES, cudaGetErrorString(e)); return (int)e; }
    const int grid = p.n_items < sms ? p.n_items : sms;
    corr_fwd_tc<T, NCHW><<<grid, THREADS, SMEM_BYTES, st>>>(tmA, tmB, p);
    return check_launch("corr_fwd_tc");
}

}  // namespace tc

int corr_tc_main(float *out, const CorrGeom &g, void *ws, size_t ws_bytes, cudaStream_t st,
                 bool nchw_out, int c_dst, int c_off, float slope)
{
    float *P1, *P2, *tmp;
    int rc = tc_ws(g, ws, ws_bytes, false, P1, P2, tmp, "corr_fwd");
    if (rc) return rc;
    if (!nchw_out) return tc::tc_launch<tc::Tile<16, 8>, false>(out, g, P1, P2, st, c_dst, c_off, slope);
    // NCHW: the tile shape that wastes fewer partial tiles (16 x 8 on a tie)
    const int PH = g.H / 2, PW = g.W / 2;
    const long long t168 = (long long)((PH + 15) / 16) * ((PW + 7) / 8), t816 = (long long)((PH + 7) / 8) * ((PW + 15) / 16);
    if (t816 < t168) return tc::tc_launch<tc::Tile<8, 16>, true>(out, g, P1, P2, st, 0, 0, 1.f);
    return tc::tc_launch<tc::Tile<16, 8>, true>(out, g, P1, P2, st, 0, 0, 1.f);
}

}  // namespace flowops

extern "C" int flowops_corr_set_impl(int flags)
{
    flowops::g_corr_impl = flags;
    return 0;
}

// Debugging aid (not part of the operator API): device buffer of 8 x 148 uint64 that the next tensor-core launches fill
// with per-CTA cycle counters -- [0] TMA producer waiting for a free stage, [1] UMMA thread waiting for split data,
// [2] UMMA thread waiting for the epilogue to drain TMEM, [3] splitters waiting for TMA data, [4] epilogue waiting for the
// accumulator, [5] epilogue busy, [6] CTA lifetime, [7] splitters busy.  NULL switches it off.
extern "C" int flowops_corr_tc_trace(void *device_buffer)
{
    flowops::g_tc_trace = static_cast<unsigned long long *>(device_buffer);
    return 0;
}

extern "C" int flowops_corr_get_impl(void) { return flowops::corr_impl_flags(); }
