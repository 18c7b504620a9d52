// bf16 tensor-core GEMM for sm_100a: TMA (128B swizzle) -> shared-memory ring -> tcgen05.mma (fp32 accumulators
// in TMEM) -> tcgen05.ld epilogue.  Warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (single thread),
// warps 2..5 = epilogue (one TMEM lane quadrant each).  CTA tile 128 x 128 x 64, up to 4 stages.  Split-K writes one
// partial slab per split (plain stores) and a small kernel adds the slabs in split order: no float atomics anywhere,
// so the same inputs give the same bits on every run.  Both operands may be K-major or MN-major (the dW = dG^T Z products of BPTT contract over
// the row index of both stored matrices).
//
// D[M,N] = op(A)[M,K] * op(B)[K,N] (+ bias[N]);  A: transA=0 stored [M,K] (K-major), transA=1 stored [K,M]
// (MN-major); B: transB=1 stored [N,K] (K-major), transB=0 stored [K,N] (MN-major).
#include <cuda.h>
#include <mutex>

#include "tc.cuh"
#include "gemm_tc.cuh"

namespace csn {

using namespace tc;

constexpr int GBM = 128, GBK = 64;  // the N tile (GBN) is a template parameter: 128, or 32 for the per-timestep LSTM-cell GEMM
constexpr int kGemmThreads = 192;
template <int GBN>
constexpr uint32_t stage_bytes() { return uint32_t(GBM * GBK + GBN * GBK) * 2u; }  // 32 KB at GBN = 128
constexpr int kEpiStride = 36;                                  // floats; 16-byte aligned rows, conflict-free float4
constexpr uint32_t kEpiBytes = 4 * 32 * kEpiStride * 4;         // one 32x32 staging tile per epilogue warp

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D bf16 tensor [outer, inner] with row pitch `pitch_elems`, box [box_outer, box_inner], 128B swizzle.
int make_tmap_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                        uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return CSN_ECUDA;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu pitch=%llu)", (int)r,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_elems);
    return CSN_ECUDA;
  }
  return CSN_OK;
}

int make_tmap_2d_plain(CUtensorMap* tm, const void* base, int elem_bytes, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                       uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return CSN_ECUDA;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base),
                  dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (plain) failed with CUresult %d (inner=%llu outer=%llu pitch=%llu box=%u x %u)", (int)r,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_elems, box_inner, box_outer);
    return CSN_ECUDA;
  }
  return CSN_OK;
}

__device__ __forceinline__ float tanh_fast_g(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast_g(float x) { return fmaf(0.5f, tanh_fast_g(0.5f * x), 0.5f); }

template <bool A_MN, bool B_MN, int GBN>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                  const __grid_constant__ CUtensorMap tmB,
                                                                  const GemmEpi p) {
  static_assert(GBN == 128 || (GBN == 32 && !B_MN), "N tile: 128, or 32 with a K-major B operand");
  constexpr uint32_t kStageBytes = stage_bytes<GBN>();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + size_t(S) * kStageBytes);
  uint64_t* empty = full + 4;
  uint64_t* accfull = empty + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
  const int kbeg = blockIdx.z * p.k_per_split;
  const int klen = min(p.K - kbeg, p.k_per_split);
  const int n_it = p.zero_acc ? 0 : (klen + GBK - 1) / GBK;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(accfull, 1);
    fence_mbar_init();
  }
  if (p.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the next step's grid may start its prologue
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmB); }
  if (warp == 1) tmem_alloc(tmem_slot, GBN);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above touched no global data; from here on the predecessor grid's results are needed
  if (p.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < n_it; ++it) {
        const int s = it % S;
        if (it >= S) mbar_wait(&empty[s], ((it / S) - 1) & 1);
        uint8_t* sa = smem + size_t(s) * kStageBytes;
        uint8_t* sb = sa + GBM * GBK * 2;
        mbar_arrive_expect_tx(&full[s], kStageBytes);
        const int k = kbeg + it * GBK;
        if (A_MN) {
          tma_load_2d(sa, &tmA, &full[s], m0, k);
          tma_load_2d(sa + 8192, &tmA, &full[s], m0 + 64, k);
        } else {
          tma_load_2d(sa, &tmA, &full[s], k, m0);
        }
        if (B_MN) {
          tma_load_2d(sb, &tmB, &full[s], n0, k);
          tma_load_2d(sb + 8192, &tmB, &full[s], n0 + 64, k);
        } else {
          tma_load_2d(sb, &tmB, &full[s], k, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(GBM, GBN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      for (int it = 0; it < n_it; ++it) {
        const int s = it % S;
        mbar_wait(&full[s], (it / S) & 1);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + size_t(s) * kStageBytes);
        const uint32_t sb = sa + GBM * GBK * 2;
#pragma unroll
        for (int kk = 0; kk < GBK / 16; ++kk) {
          // K-major SW128: rows of 128 B, 8-row groups 1024 B apart, 16 K-elements = 32 B along the row.
          // MN-major SW128: 8 k-rows x 128 B atoms 1024 B apart along K (SBO), 64-element MN chunks 8192 B apart (LBO).
          uint64_t da = A_MN ? make_smem_desc(sa + kk * 2048, 8192, 1024, kLayoutSw128)
                             : make_smem_desc(sa + kk * 32, 16, 1024, kLayoutSw128);
          uint64_t db = B_MN ? make_smem_desc(sb + kk * 2048, 8192, 1024, kLayoutSw128)
                             : make_smem_desc(sb + kk * 32, 16, 1024, kLayoutSw128);
          umma_f16(tmem_base, da, db, idesc, (it | kk) != 0);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(accfull);
    }
  } else {
    // epilogue: warp w may only touch TMEM lanes [32*(w%4), +32).  Each warp drains its 32 rows in 32-column
    // chunks through a private padded shared tile so that global stores are row-contiguous (8 lanes x float4 = one
    // 128-byte line per row) instead of one 4-byte store per lane per row.
    const int quad = warp & 3;
    float* tile = reinterpret_cast<float*>(smem + size_t(S) * kStageBytes + 256) + quad * (32 * kEpiStride);
    if (p.mode == 1) {
      // ---- fused LSTM cell: (i,f,g,o) = act(acc + Xp); c = f c_prev + i g; h = o tanh(c) ----
      // The cell's other inputs (Xp float4, c_prev) do not depend on the MMAs: they are fetched one 32-column chunk
      // AHEAD into registers (the first chunk before the accumulator wait), so their HBM/L2 latency hides behind the
      // TMA/MMA pipeline instead of being paid eight dependent times per chunk (the stores of one cell may alias the
      // loads of the next as far as the compiler can tell, so it cannot hoist them itself).
      const int col4 = (lane & 7) * 4;
      constexpr int kChunks = GBN / 32;
      auto fetch = [&](int c, float4 (&x4)[8], float (&cp)[8]) {
        const int gn = n0 + c * 32 + col4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int gm = m0 + quad * 32 + i * 4 + (lane >> 3);
          const bool ok = gm < p.M && gn < p.N;
          x4[i] = ok ? __ldg(reinterpret_cast<const float4*>(p.xp + size_t(gm) * p.N + gn)) : make_float4(0.f, 0.f, 0.f, 0.f);
          cp[i] = (ok && p.c_prev) ? __ldg(p.c_prev + size_t(gm) * p.H + (gn >> 2)) : 0.f;
        }
      };
      float4 xa[8];
      float ca[8];
      fetch(0, xa, ca);
      if (n_it > 0) {
        mbar_wait(accfull, 0);
        tcgen05_fence_after();
      }
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        float4 xb[8];
        float cb[8];
        const bool more = (c + 1 < kChunks) && (n0 + (c + 1) * 32 < p.N);
        if (more) fetch(c + 1, xb, cb);
        uint32_t r[32];
        if (n_it > 0) {
          tmem_ld<32>(tmem_base + (uint32_t(quad * 32) << 16) + c * 32, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4)
          *reinterpret_cast<uint4*>(tile + lane * kEpiStride + q4 * 4) = make_uint4(r[q4 * 4], r[q4 * 4 + 1], r[q4 * 4 + 2], r[q4 * 4 + 3]);
        __syncwarp();
        const int gn = n0 + c * 32 + col4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + (lane >> 3);
          const int gm = m0 + quad * 32 + rr;
          if (gm >= p.M || gn >= p.N) continue;
          const float4 v = *reinterpret_cast<const float4*>(tile + rr * kEpiStride + col4);
          const int u = gn >> 2;
          const float ig = sigmoid_fast_g(v.x + xa[i].x), fg = sigmoid_fast_g(v.y + xa[i].y);
          const float gg = tanh_fast_g(v.z + xa[i].z), og = sigmoid_fast_g(v.w + xa[i].w);
          const float cc = fmaf(fg, ca[i], ig * gg);
          p.c_out[size_t(gm) * p.H + u] = cc;
          p.h_out[size_t(gm) * p.H + u] = __float2bfloat16_rn(og * tanh_fast_g(cc));
          if (p.gates_out) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(ig, fg), hi = __floats2bfloat162_rn(gg, og);
            uint2 o;
            o.x = *reinterpret_cast<uint32_t*>(&lo);
            o.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(p.gates_out + size_t(gm) * p.N + gn) = o;
          }
        }
        __syncwarp();
        if (!more) break;
#pragma unroll
        for (int i = 0; i < 8; ++i) { xa[i] = xb[i]; ca[i] = cb[i]; }
      }
    } else {
    if (n_it > 0) {
      mbar_wait(accfull, 0);
      tcgen05_fence_after();
    }
    const bool f32_out = (p.d_dtype == CSN_F32);
    // split-K into per-split slabs (no atomics): split z owns D + z * split_stride elements; the consumer sums them
    void* const Dz = p.split_stride ? static_cast<void*>(reinterpret_cast<float*>(p.D) + size_t(blockIdx.z) * p.split_stride) : p.D;
    const bool vec_ok = ((p.ldd & 3) == 0) && ((reinterpret_cast<uintptr_t>(Dz) & 15) == 0);
    const bool add_bias = p.bias && blockIdx.z == 0;
#pragma unroll 1
    for (int c0 = 0; c0 < GBN; c0 += 32) {
      if (n0 + c0 >= p.N) break;  // warp-uniform
      uint32_t r[32];
      if (n_it > 0) {
        tmem_ld<32>(tmem_base + (uint32_t(quad * 32) << 16) + c0, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4)
        *reinterpret_cast<uint4*>(tile + lane * kEpiStride + q4 * 4) = make_uint4(r[q4 * 4], r[q4 * 4 + 1], r[q4 * 4 + 2], r[q4 * 4 + 3]);
      __syncwarp();
      const int col4 = (lane & 7) * 4;
      const int gn = n0 + c0 + col4;
      float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (add_bias) {
        bv.x = gn < p.N ? p.bias[gn] : 0.f;
        bv.y = gn + 1 < p.N ? p.bias[gn + 1] : 0.f;
        bv.z = gn + 2 < p.N ? p.bias[gn + 2] : 0.f;
        bv.w = gn + 3 < p.N ? p.bias[gn + 3] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = i * 4 + (lane >> 3);
        const int gm = m0 + quad * 32 + rr;
        float4 v = *reinterpret_cast<const float4*>(tile + rr * kEpiStride + col4);
        v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
        if (gm >= p.M || gn >= p.N) continue;
        if (f32_out) {
          float* d = reinterpret_cast<float*>(Dz) + size_t(gm) * p.ldd + gn;
          if (p.atomic) {  // accumulate into D: this CTA owns the tile (split_k == 1), a plain read-modify-write
            d[0] += v.x;
            if (gn + 1 < p.N) d[1] += v.y;
            if (gn + 2 < p.N) d[2] += v.z;
            if (gn + 3 < p.N) d[3] += v.w;
          } else if (vec_ok && gn + 3 < p.N) {
            *reinterpret_cast<float4*>(d) = v;
          } else {
            d[0] = v.x;
            if (gn + 1 < p.N) d[1] = v.y;
            if (gn + 2 < p.N) d[2] = v.z;
            if (gn + 3 < p.N) d[3] = v.w;
          }
        } else {
          __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(p.D) + size_t(gm) * p.ldd + gn;
          if (vec_ok && gn + 3 < p.N) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            uint2 o;
            o.x = *reinterpret_cast<uint32_t*>(&lo);
            o.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(d) = o;
          } else {
            d[0] = __float2bfloat16_rn(v.x);
            if (gn + 1 < p.N) d[1] = __float2bfloat16_rn(v.y);
            if (gn + 2 < p.N) d[2] = __float2bfloat16_rn(v.z);
            if (gn + 3 < p.N) d[3] = __float2bfloat16_rn(v.w);
          }
        }
      }
      __syncwarp();
    }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, GBN);
}

template <bool A_MN, bool B_MN, int GBN = 128>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpi& p, dim3 grid, cudaStream_t s) {
  const size_t smem = size_t(p.stages) * stage_bytes<GBN>() + 1024 + 256 + kEpiBytes;
  auto kern = gemm_tc_kernel<A_MN, B_MN, GBN>;
  static size_t smem_set = 0;  // per instantiation; raise-only, so the call disappears from steady state (and from graph capture)
  if (smem > smem_set) {
    CSN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  if (p.pdl) {
    // programmatic dependent launch: this grid may be scheduled (and run its prologue: barrier init, TMEM allocation,
    // descriptor prefetch) while its predecessor in the stream drains; griddepcontrol.wait in the kernel orders the data
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kGemmThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CSN_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, p));
    count_launches(1);
    return CSN_OK;
  }
  kern<<<grid, kGemmThreads, smem, s>>>(ta, tb, p);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent variant for plain products with MANY output tiles and a short K (the hoisted input projection of the
// H = 512 encoder: 7040 tiles of K = 128; the K = 65536 DINO head: 1536 tiles of K = 256).  One tile per CTA pays
// barrier init + TMEM allocation + the drain of a 64 KB tile for two to four k-iterations of tensor work, and nothing
// overlaps the drain (measured: 332 us for the 29.5 GFLOP / 461 MB projection).  Here a CTA walks tiles
// blockIdx.x, blockIdx.x + gridDim.x, ...: the TMA producer and the MMA issuer run ahead into the next tile through the
// same shared-memory ring, the accumulator is DOUBLE-BUFFERED in tensor memory (2 x 128 columns), and the four epilogue
// warps drain tile i while tile i + 1 multiplies.  Plain store / bias / fp32 or bf16 output only.
template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                             const __grid_constant__ CUtensorMap tmB,
                                                                             const GemmEpi p, int tiles_n, int n_tiles) {
  constexpr int GBN = 128;
  constexpr uint32_t kStageBytes = stage_bytes<GBN>();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + size_t(S) * kStageBytes);
  uint64_t* empty = full + 4;
  uint64_t* accfull = empty + 4;    // [2]
  uint64_t* accempty = accfull + 2; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_it = (p.K + GBK - 1) / GBK;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&accfull[i], 1); mbar_init(&accempty[i], 4); }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmB); }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * GBN);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t git = 0;  // ring position, continuous over the tiles
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * GBM, n0 = (tile % tiles_n) * GBN;
        for (int it = 0; it < n_it; ++it, ++git) {
          const uint32_t s = git % S;
          if (git >= (uint32_t)S) mbar_wait(&empty[s], ((git / S) - 1) & 1);
          uint8_t* sa = smem + size_t(s) * kStageBytes;
          uint8_t* sb = sa + GBM * GBK * 2;
          mbar_arrive_expect_tx(&full[s], kStageBytes);
          const int k = it * GBK;
          if (A_MN) {
            tma_load_2d(sa, &tmA, &full[s], m0, k);
            tma_load_2d(sa + 8192, &tmA, &full[s], m0 + 64, k);
          } else {
            tma_load_2d(sa, &tmA, &full[s], k, m0);
          }
          if (B_MN) {
            tma_load_2d(sb, &tmB, &full[s], n0, k);
            tma_load_2d(sb + 8192, &tmB, &full[s], n0 + 64, k);
          } else {
            tma_load_2d(sb, &tmB, &full[s], k, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(GBM, GBN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      uint32_t git = 0, lt = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
        const uint32_t a = lt & 1;
        if (lt >= 2) mbar_wait(&accempty[a], ((lt >> 1) - 1) & 1);  // the epilogue has drained this buffer's previous tile
        tcgen05_fence_after();
        for (int it = 0; it < n_it; ++it, ++git) {
          const uint32_t s = git % S;
          mbar_wait(&full[s], (git / S) & 1);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + size_t(s) * kStageBytes);
          const uint32_t sb = sa + GBM * GBK * 2;
#pragma unroll
          for (int kk = 0; kk < GBK / 16; ++kk) {
            uint64_t da = A_MN ? make_smem_desc(sa + kk * 2048, 8192, 1024, kLayoutSw128)
                               : make_smem_desc(sa + kk * 32, 16, 1024, kLayoutSw128);
            uint64_t db = B_MN ? make_smem_desc(sb + kk * 2048, 8192, 1024, kLayoutSw128)
                               : make_smem_desc(sb + kk * 32, 16, 1024, kLayoutSw128);
            umma_f16(tmem_base + a * GBN, da, db, idesc, (it | kk) != 0);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&accfull[a]);
      }
    }
  } else {
    const int quad = warp & 3;
    float* tile_s = reinterpret_cast<float*>(smem + size_t(S) * kStageBytes + 256) + quad * (32 * kEpiStride);
    const bool f32_out = (p.d_dtype == CSN_F32);
    const bool vec_ok = ((p.ldd & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.D) & 15) == 0);
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
      const int m0 = (tile / tiles_n) * GBM, n0 = (tile % tiles_n) * GBN;
      const uint32_t a = lt & 1;
      mbar_wait(&accfull[a], (lt >> 1) & 1);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < GBN; c0 += 32) {
        if (n0 + c0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld<32>(tmem_base + (uint32_t(quad * 32) << 16) + a * GBN + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4)
          *reinterpret_cast<uint4*>(tile_s + lane * kEpiStride + q4 * 4) = make_uint4(r[q4 * 4], r[q4 * 4 + 1], r[q4 * 4 + 2], r[q4 * 4 + 3]);
        __syncwarp();
        const int col4 = (lane & 7) * 4;
        const int gn = n0 + c0 + col4;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) {
          bv.x = gn < p.N ? p.bias[gn] : 0.f;
          bv.y = gn + 1 < p.N ? p.bias[gn + 1] : 0.f;
          bv.z = gn + 2 < p.N ? p.bias[gn + 2] : 0.f;
          bv.w = gn + 3 < p.N ? p.bias[gn + 3] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + (lane >> 3);
          const int gm = m0 + quad * 32 + rr;
          float4 v = *reinterpret_cast<const float4*>(tile_s + rr * kEpiStride + col4);
          v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
          if (gm >= p.M || gn >= p.N) continue;
          if (f32_out) {
            float* d = reinterpret_cast<float*>(p.D) + size_t(gm) * p.ldd + gn;
            if (vec_ok && gn + 3 < p.N) {
              *reinterpret_cast<float4*>(d) = v;
            } else {
              d[0] = v.x;
              if (gn + 1 < p.N) d[1] = v.y;
              if (gn + 2 < p.N) d[2] = v.z;
              if (gn + 3 < p.N) d[3] = v.w;
            }
          } else {
            __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(p.D) + size_t(gm) * p.ldd + gn;
            if (vec_ok && gn + 3 < p.N) {
              __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
              uint2 o;
              o.x = *reinterpret_cast<uint32_t*>(&lo);
              o.y = *reinterpret_cast<uint32_t*>(&hi);
              *reinterpret_cast<uint2*>(d) = o;
            } else {
              d[0] = __float2bfloat16_rn(v.x);
              if (gn + 1 < p.N) d[1] = __float2bfloat16_rn(v.y);
              if (gn + 2 < p.N) d[2] = __float2bfloat16_rn(v.z);
              if (gn + 3 < p.N) d[3] = __float2bfloat16_rn(v.w);
            }
          }
        }
        __syncwarp();
      }
      tcgen05_fence_before();  // this warp's TMEM reads are ordered before the MMAs the arrive below releases
      __syncwarp();
      if (lane == 0) mbar_arrive(&accempty[a]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * GBN);
}

template <bool A_MN, bool B_MN>
static int launch_gemm_persistent(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpi& p, int tiles_m, int tiles_n,
                                  cudaStream_t s) {
  const size_t smem = size_t(p.stages) * stage_bytes<128>() + 1024 + 256 + kEpiBytes;
  auto kern = gemm_tc_persistent_kernel<A_MN, B_MN>;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    CSN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  const int n_tiles = tiles_m * tiles_n;
  const int grid = std::min(n_tiles, sm_count());
  kern<<<grid, kGemmThreads, smem, s>>>(ta, tb, p, tiles_n, n_tiles);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}

// D[m, n] (+)= sum_z slabs[z][m, n] (+ bias[n]), z ascending: the fixed-order tail of a split-K product
__global__ void splitk_reduce_kernel(const float* __restrict__ slabs, size_t stride, int S, float* __restrict__ D, int ldd,
                                     int M, int N, const float* __restrict__ bias, int accumulate) {
  const size_t total = size_t(M) * N;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int m = int(i / N), n = int(i - size_t(m) * N);
    float a = slabs[i];
    for (int z = 1; z < S; ++z) a += slabs[size_t(z) * stride + i];
    if (bias) a += bias[n];
    float* d = D + size_t(m) * ldd + n;
    *d = accumulate ? *d + a : a;
  }
}

int gemm_tc_splits(int K, int split_k) {
  const int k_iters = ceil_div(K, GBK);
  split_k = std::max(1, std::min(split_k, k_iters));
  return ceil_div(k_iters, ceil_div(k_iters, split_k));
}

int splitk_reduce(const float* slabs, size_t stride, int S, float* D, int ldd, int M, int N, const float* bias,
                  int accumulate, cudaStream_t s) {
  const size_t total = size_t(M) * N;
  const int blocks = (int)std::min<size_t>(ceil_div<size_t>(total, 256), size_t(sm_count()) * 8);
  splitk_reduce_kernel<<<blocks, 256, 0, s>>>(slabs, stride, S, D, ldd, M, N, bias, accumulate);
  CSN_LAUNCH_CHECK();
  return CSN_OK;
}

}  // namespace csn

using namespace csn;

extern "C" int csn_gemm_bf16_tc(int transA, int transB, int M, int N, int K, const void* A, int lda, const void* B,
                                int ldb, void* D, int ldd, int d_dtype, const float* bias, int accumulate, int split_k,
                                void* workspace, void* stream) {
  cudaStream_t s = as_stream(stream);
  if (split_k < 1) split_k = 1;
  split_k = std::min(split_k, ceil_div(K, GBK));
  if (split_k == 1) return gemm_tc_run(transA, transB, M, N, K, A, lda, B, ldb, D, ldd, d_dtype, bias, accumulate, 1, nullptr, s);
  CSN_REQUIRE(workspace, "csn_gemm_bf16_tc: split_k > 1 needs a workspace of split_k * M * N floats (partial slabs)");
  CSN_REQUIRE(d_dtype == CSN_F32 && D, "csn_gemm_bf16_tc: split-K needs an fp32 output");
  GemmEpi slab{};
  slab.split_stride = size_t(M) * N;
  CSN_TRY(gemm_tc_run(transA, transB, M, N, K, A, lda, B, ldb, workspace, N, CSN_F32, nullptr, 0, split_k, &slab, s));
  return splitk_reduce(reinterpret_cast<const float*>(workspace), slab.split_stride, gemm_tc_splits(K, split_k),
                       reinterpret_cast<float*>(D), ldd, M, N, bias, accumulate, s);
}

int csn::gemm_tc_run(int transA, int transB, int M, int N, int K, const void* A, int lda, const void* B, int ldb, void* D,
                     int ldd, int d_dtype, const float* bias, int accumulate, int split_k, const GemmEpi* cell,
                     cudaStream_t s) {
  CSN_REQUIRE(A && B && (D || cell), "csn_gemm_bf16_tc: null pointer");
  CSN_REQUIRE(M >= 1 && N >= 1 && K >= 1, "csn_gemm_bf16_tc: dimensions must be positive");
  CSN_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "csn_gemm_bf16_tc: lda/ldb must be multiples of 8 elements (TMA 16 B pitch)");
  CSN_REQUIRE(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0,
              "csn_gemm_bf16_tc: operands must be 16-byte aligned");
  CSN_REQUIRE(d_dtype == CSN_F32 || d_dtype == CSN_BF16, "csn_gemm_bf16_tc: bad d_dtype");
  if (split_k < 1) split_k = 1;
  CSN_REQUIRE(d_dtype == CSN_F32 || (split_k == 1 && !accumulate), "csn_gemm_bf16_tc: bf16 output cannot accumulate / split-K");
  if (cell && !cell->split_stride) {
    CSN_REQUIRE(split_k == 1 && !accumulate && N % 4 == 0 && cell->xp && cell->h_out && cell->c_out && cell->H * 4 == N,
                "gemm_tc_run: bad LSTM-cell epilogue arguments");
  }

  const int k_iters = ceil_div(K, GBK);
  if (split_k > k_iters) split_k = k_iters;
  const int it_per_split = ceil_div(k_iters, split_k);
  split_k = ceil_div(k_iters, it_per_split);

  CUtensorMap ta, tb;
  if (transA) CSN_TRY(make_tmap_2d(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, 64));   // stored [K, M]
  else        CSN_TRY(make_tmap_2d(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, 128));  // stored [M, K]
  // LSTM-cell epilogue on a small batch: 32-column tiles put 4x more CTAs on each step of the serial chain
  const bool narrow = cell && transB && !transA && ceil_div(N, 128) * ceil_div(M, GBM) * 4 <= 2 * sm_count();
  const int gbn = narrow ? 32 : 128;
  if (transB) CSN_TRY(make_tmap_2d(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, gbn));  // stored [N, K]
  else        CSN_TRY(make_tmap_2d(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, 64));   // stored [K, N]

  GemmEpi p{};
  p.D = D; p.bias = bias; p.ldd = ldd; p.d_dtype = d_dtype; p.M = M; p.N = N; p.K = K;
  p.k_per_split = it_per_split * GBK;
  p.split_stride = 0;
  p.pdl = cell ? cell->pdl : 0;
  if (cell && cell->split_stride) {  // caller-provided slabs: plain stores, split z -> D + z * split_stride
    CSN_REQUIRE(!cell->xp && d_dtype == CSN_F32 && !accumulate, "gemm_tc_run: split slabs need fp32 output without accumulation");
    p.split_stride = cell->split_stride;
  }
  CSN_REQUIRE(split_k == 1 || p.split_stride, "gemm_tc_run: split-K needs partial slabs (GemmEpi::split_stride)");
  p.atomic = (!p.split_stride && accumulate) ? 1 : 0;  // read-modify-write epilogue
  p.stages = it_per_split < 4 ? (it_per_split < 2 ? 2 : it_per_split) : 4;
  if (cell && !cell->split_stride) {
    p.mode = 1; p.zero_acc = cell->zero_acc; p.H = cell->H; p.xp = cell->xp; p.c_prev = cell->c_prev;
    p.h_out = cell->h_out; p.gates_out = cell->gates_out; p.c_out = cell->c_out;
  }
  dim3 grid(ceil_div(N, gbn), ceil_div(M, GBM), split_k);
  {
    // many output tiles, each with little tensor work: walk them with persistent CTAs (double-buffered accumulators)
    static const bool no_persist = [] { const char* e = getenv("CSN_GEMM_NO_PERSISTENT"); return e && e[0] == '1'; }();
    const long long tiles = (long long)grid.x * grid.y;
    if (!no_persist && !cell && split_k == 1 && !accumulate && gbn == 128 && k_iters <= 64 && tiles >= sm_count() + sm_count() / 4 &&
        tiles < (1ll << 30)) {
      p.stages = 4;
      if (transA && !transB) return launch_gemm_persistent<true, true>(ta, tb, p, (int)grid.y, (int)grid.x, s);
      if (transA && transB) return launch_gemm_persistent<true, false>(ta, tb, p, (int)grid.y, (int)grid.x, s);
      if (!transA && !transB) return launch_gemm_persistent<false, true>(ta, tb, p, (int)grid.y, (int)grid.x, s);
      return launch_gemm_persistent<false, false>(ta, tb, p, (int)grid.y, (int)grid.x, s);
    }
  }
  CSN_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "csn_gemm_bf16_tc: grid too large");
  if (transA && !transB) return launch_gemm<true, true>(ta, tb, p, grid, s);
  if (transA && transB) return launch_gemm<true, false>(ta, tb, p, grid, s);
  if (!transA && !transB) return launch_gemm<false, true>(ta, tb, p, grid, s);
  if (narrow) return launch_gemm<false, false, 32>(ta, tb, p, grid, s);
  return launch_gemm<false, false>(ta, tb, p, grid, s);
}
