// Internal interface of the tcgen05 GEMM (gemm_tc.cu) for the other translation units of libcsn_b200.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace csn {

struct GemmEpi {
  void* D;
  const float* bias;
  int ldd, d_dtype, M, N, K, k_per_split, atomic /* accumulate into D (read-modify-write) */, stages;
  // split-K without atomics: when non-zero, split z stores its partial product to D + z * split_stride (fp32 elements);
  // the consumer adds the slabs (lstm_cell_bwd_kernel does, lstm_tc_large.cu).  Set in a `cell` argument with xp == NULL.
  size_t split_stride;
  int pdl;  // launch with programmatic stream serialization (the per-timestep chains of lstm_tc_large.cu)
  // mode 1: fused LSTM cell epilogue (large-hidden path, see lstm_tc_large.cu).  Columns are gate-interleaved
  // (column 4u+g = gate g of hidden unit u), so the float4 a lane owns is one cell's (i,f,g,o) pre-activation.
  int mode, zero_acc, H;
  const float* xp;            // [M, N] fp32: hoisted input projection (+ biases) of this timestep
  const float* c_prev;        // [M, H] fp32 or NULL (t = 0)
  __nv_bfloat16* h_out;       // [M, H]
  __nv_bfloat16* gates_out;   // [M, N] activated gates (BPTT reserve) or NULL
  float* c_out;               // [M, H]
};

// Same contract as csn_gemm_bf16_tc; `cell` (may be NULL) switches the epilogue to the fused LSTM cell.
int gemm_tc_run(int transA, int transB, int M, int N, int K, const void* A, int lda, const void* B, int ldb, void* D,
                int ldd, int d_dtype, const float* bias, int accumulate, int split_k, const GemmEpi* cell, cudaStream_t s);

// TMA descriptor of a 2-D bf16 tensor [outer, inner] (row pitch `pitch_elems`), box [box_outer, box_inner], 128B swizzle
int make_tmap_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_elems, uint32_t box_inner,
                 uint32_t box_outer);

// TMA descriptor of a plain (unswizzled) 2-D tensor of 2- or 4-byte elements (elem_bytes: 2 = bf16, 4 = fp32)
int make_tmap_2d_plain(CUtensorMap* tm, const void* base, int elem_bytes, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                       uint32_t box_inner, uint32_t box_outer);

// D[m, n] (+)= sum_z slabs[z * stride + m * N + n] (+ bias[n]), z ascending (the fixed-order tail of a split-K product)
int splitk_reduce(const float* slabs, size_t stride, int S, float* D, int ldd, int M, int N, const float* bias,
                  int accumulate, cudaStream_t s);
// number of splits gemm_tc_run actually launches for a requested split_k
int gemm_tc_splits(int K, int split_k);

}  // namespace csn
