// api.cu -- error plumbing, version, and the FFMA-peak measurement helper of libflowops.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace flowops {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Launch-time errors only (cudaPeekAtLastError does not synchronise, so calls stay async and
// CUDA-graph capturable).  The reference does the same with cudaGetLastError
// (correlation_cuda_kernel.cu:417-424) and turns it into a RuntimeError; the Python layer here
// does likewise.
int check_launch(const char *what)
{
    // peek, do not clear: a pending error left by an earlier launch of the host framework stays visible to it
    cudaError_t e = cudaPeekAtLastError();
    // FLOWOPS_DEBUG_SYNC=1: synchronise after every launch so that an execution error is attributed
    // to the kernel that caused it (development aid; never set in production or under graph capture)
    static const bool debug_sync = getenv("FLOWOPS_DEBUG_SYNC") != nullptr;
    if (e == cudaSuccess && debug_sync) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        // reported to the caller here, so take it off the runtime's error state (non-sticky errors would otherwise be
        // reported again by the next call); it may stem from an earlier launch on this thread, hence the wording
        (void)cudaGetLastError();
        set_error("%s: %s (reported after this launch; a pending error of an earlier launch is reported here too)", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

// FP32-FMA pipe peak: 64 independent accumulator chains per thread, acc = fma(acc, x, y) with x, y
// loop-invariant, i.e. one register-file operand per FFMA -- the pattern that is limited by the FMA pipe
// itself and not by register-bank conflicts (tools/ffma_probe.cu: 72.5 TFLOP/s on B200 vs 66 for an
// 8x8 outer product and 40 for a naive sliding-window tile).  8 warps x 2 CTAs per SM.
__global__ void __launch_bounds__(256, 2) ffma_peak_kernel(float *sink, int iters, float a, float b)
{
    float acc[64];
    const float x = a + threadIdx.x * 1e-7f, y = b;
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = (float)(i + threadIdx.x);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 64; ++i) acc[i] = __fmaf_rn(acc[i], x, y);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) s += acc[i];
    if (s == 123.456f) *sink = s;   // never true in practice; keeps the chains alive
}

}  // namespace flowops

using namespace flowops;

extern "C" int flowops_version(void) { return FLOWOPS_VERSION; }

extern "C" const char *flowops_last_error(void) { return g_err; }

extern "C" int flowops_bench_ffma(float *sink, int iters, double *flops, void *stream)
{
    FLOWOPS_REQUIRE(sink && iters > 0, FLOWOPS_EINVAL, "bench_ffma: bad arguments");
    const int grid = kNumSMs * 2;
    ffma_peak_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(sink, iters, 0.999f, 0.001f);
    if (flops) *flops = 2.0 * 64.0 * (double)iters * 256.0 * (double)grid;
    return check_launch("bench_ffma");
}
