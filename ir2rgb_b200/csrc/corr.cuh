// corr.cuh -- geometry shared by the Correlation kernels.
#pragma once
#include "common.cuh"

namespace flowops {

struct CorrGeom {
    int B, C, H, W;          // inputs  [B,C,H,W]
    int pad, k, md, s1, s2;  // reference parameters (correlation.py:43)
    int kr, dr, D;           // kernel radius, displacement radius (md / s2), D = 2*dr + 1
    int oC, oH, oW;          // output [B,oC,oH,oW], correlation_cuda.cc:19-34
};

// Fills g; returns 0 or a FLOWOPS_E* code (message set).
int corr_geometry(CorrGeom &g, int B, int C, int H, int W, int pad, int k, int md, int s1, int s2);

// FlowNetC configuration the fast path is specialised for (FlowNetC.py:31):
// k = 1, s1 = 1, s2 = 2, pad == md == 20  (D = 21, 441 output channels).
bool corr_fast_supported(const CorrGeom &g);

int corr_fwd_generic_launch(const float *in1, const float *in2, float *out, const CorrGeom &g, cudaStream_t st);
int corr_bwd_generic_launch(const float *in1, const float *in2, const float *gout, float *gin1, float *gin2,
                            const CorrGeom &g, cudaStream_t st);

size_t corr_fast_fwd_workspace(const CorrGeom &g);
size_t corr_fast_bwd_workspace(const CorrGeom &g);
int corr_fast_fwd_launch(const float *in1, const float *in2, float *out, const CorrGeom &g, int in_layout,
                         void *ws, size_t ws_bytes, cudaStream_t st, int io_dtype = 0);
int corr_fast_planes_nhwc(const float *in1, const float *in2, const CorrGeom &g, int only, const float *bias, float slope,
                          float *act, void *ws, size_t ws_bytes, cudaStream_t st);
int corr_fast_main(float *out, const CorrGeom &g, void *ws, size_t ws_bytes, cudaStream_t st,
                   bool nhwc_out = false, int c_dst = 0, int c_off = 0, float slope = 1.f, int out_dtype = 0);
int corr_fast_bwd_launch(const float *in1, const float *in2, const float *gout, float *gin1, float *gin2,
                         const CorrGeom &g, void *ws, size_t ws_bytes, cudaStream_t st);

// Tensor-core (tcgen05, 3xTF32) forward for the FlowNetC configuration: csrc/corr_tc.cu.  Selected when supported
// unless FLOWOPS_CORR_IMPL=ffma (or flowops_corr_set_impl(0)); the FP32-FMA kernel above stays the general path.
int corr_impl_flags();
unsigned long long *corr_tc_trace_buffer();     // debugging: see flowops_corr_tc_trace
bool corr_tc_supported(const CorrGeom &g);
size_t corr_tc_fwd_workspace(const CorrGeom &g, bool nchw_out);
int corr_tc_planes_nchw(const float *in1, const float *in2, const CorrGeom &g, void *ws, size_t ws_bytes, cudaStream_t st);
int corr_tc_planes_nhwc(const float *in, int which, const CorrGeom &g, const float *bias, float slope, float *act,
                        void *ws, size_t ws_bytes, cudaStream_t st);
int corr_tc_main(float *out, const CorrGeom &g, void *ws, size_t ws_bytes, cudaStream_t st,
                 bool nchw_out, int c_dst, int c_off, float slope);

// Tensor-core backward (csrc/corr_tc_bwd.cu): both gradients in one persistent launch after two layout passes.
bool corr_tc_bwd_supported(const CorrGeom &g);
size_t corr_tc_bwd_workspace(const CorrGeom &g);
int corr_tc_bwd_launch(const float *in1, const float *in2, const float *gout, float *gin1, float *gin2,
                       const CorrGeom &g, void *ws, size_t ws_bytes, cudaStream_t st);

}  // namespace flowops
