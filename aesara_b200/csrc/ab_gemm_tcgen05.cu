// ab_gemm_tcgen05.cu — float32 Gemm/Dot22 on the 5th-generation tensor cores.
//
// Reference op: aesara/tensor/blas.py:872 Gemm / :1659 Dot22 (C template
// :518-869 dispatching to sgemm_ by stride case :765-776).  The reference is a
// true-fp32 sgemm, so the default mode here is fp32-faithful:
//
//   precision 0  3xTF32: each operand x is split into hi = tf32(x) and
//                lo = tf32(x - hi); D += Ahi*Bhi + Ahi*Blo + Alo*Bhi with FP32
//                accumulation in TMEM (error ~2^-21 relative, rtol 1e-5 holds);
//   precision 1  one TF32 pass straight from the fp32 operands;
//   precision 2  BF16 operands, FP32 accumulation (the "bf16 compute policy"
//                of BASELINE config 3; stated looser rtol).
//
// Structure (one 128 x BLOCK_N output tile per CTA, 192 threads):
//   warp 0     TMA producer: cp.async.bulk.tensor 2-D loads of K-major,
//              128-byte-swizzled operand tiles into a multi-stage smem ring,
//              completion on mbarriers (expect_tx);
//   warp 1     allocates TMEM, then one elected lane issues tcgen05.mma
//              (cta_group::1, M=128, N=BLOCK_N, K=32 bytes per instruction);
//              tcgen05.commit releases smem stages / signals the epilogue;
//   warps 2-5  epilogue: tcgen05.ld 32x32b the accumulator lanes they own,
//              apply alpha/beta and store C with arbitrary strides.
//
// Operands that are not K-major / 16-byte-pitched in global memory (the
// DimShuffle{1,0} views of the MLP backward pass, blas.py:719-726 "unit" cases)
// and every operand of modes 0 and 2 go through pack_kernel first, which
// writes the K-major hi/lo (or bf16) planes the tensor maps describe.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>

#include "ab_common.h"

namespace ab {

namespace {

constexpr int BLOCK_M = 128;
constexpr int SW_BYTES = 128;  // swizzle span = smem row pitch of every operand tile
constexpr int kThreads = 192;
constexpr int kMaxSmem = 200 * 1024;

// ----------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
template <int KIND>  // 0 = tf32, 1 = f16/bf16
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                     uint32_t idesc, uint32_t accumulate) {
  if (KIND == 0) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute UMMA::SmemDescriptor
// field layout): start address >>4 in [0,14), LBO>>4 in [16,30), SBO>>4 in
// [32,46), version=1 in [46,48), layout type SWIZZLE_128B=2 in [61,64).
// Rows are 128 bytes apart, groups of 8 rows 1024 bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;             // LBO (unused for swizzled K-major) = 16 B
  d |= (uint64_t)(1024 >> 4) << 32;   // SBO = 8 rows * 128 B
  d |= (uint64_t)1 << 46;             // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;             // SWIZZLE_128B
  return d;
}

struct GemmParams {
  long long M, N, K;      // K in elements of the packed type
  float alpha, beta;
  float* C;
  long long c_rs, c_cs;
  int block_n;            // 64 / 128 / 256
  int acc_stages;         // TMEM accumulator stages (2 -> epilogue overlaps the next tile)
  int stages;
  int nparts;             // 1, or 2 for the hi/lo split (3 MMAs per k-step)
  int k_elems_per_row;    // elements per 128-byte smem row: 32 (tf32) or 64 (bf16)
  int a_tile_bytes, b_tile_bytes;
  uint32_t idesc;
};

template <int KIND>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a0,
                    const __grid_constant__ CUtensorMap map_a1,
                    const __grid_constant__ CUtensorMap map_b0,
                    const __grid_constant__ CUtensorMap map_b1,
                    const __grid_constant__ GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem base is only guaranteed 16-byte aligned: round up to 1024 for SWIZZLE_128B
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  __shared__ __align__(8) uint64_t full_bar[8];
  __shared__ __align__(8) uint64_t empty_bar[8];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stage_bytes = p.nparts * (p.a_tile_bytes + p.b_tile_bytes);
  const int num_k_blocks = (int)((p.K + p.k_elems_per_row - 1) / p.k_elems_per_row);
  // persistent tile scheduler: CTA b handles tiles b, b + gridDim.x, ...; N-tiles are
  // consecutive so the CTAs resident at one time share A row panels and all of B in L2
  const long long tiles_n = (p.N + p.block_n - 1) / p.block_n;
  const long long num_tiles = ((p.M + BLOCK_M - 1) / BLOCK_M) * tiles_n;
  const uint32_t tmem_cols = (uint32_t)(p.acc_stages * p.block_n);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 128);  // every epilogue thread arrives
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    // allocate the accumulator columns (power of two >= 32): acc_stages x block_n
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_slot)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a0)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b0)) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (int)((tile / tiles_n) * BLOCK_M);
        const int n0 = (int)((tile % tiles_n) * p.block_n);
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sbase = smem + (size_t)stage * stage_bytes;
          mbar_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
          const int kc = kb * p.k_elems_per_row;
          tma_load_2d(sbase, &map_a0, &full_bar[stage], kc, m0);
          tma_load_2d(sbase + p.nparts * p.a_tile_bytes, &map_b0, &full_bar[stage], kc, n0);
          if (p.nparts == 2) {
            tma_load_2d(sbase + p.a_tile_bytes, &map_a1, &full_bar[stage], kc, m0);
            tma_load_2d(sbase + 2 * p.a_tile_bytes + p.b_tile_bytes, &map_b1, &full_bar[stage], kc,
                        n0);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;  // tiles processed by this CTA
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const uint32_t as = it % (uint32_t)p.acc_stages;
        const uint32_t aphase = (it / (uint32_t)p.acc_stages) & 1u;
        // wait until the epilogue has drained this accumulator stage
        mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + as * (uint32_t)p.block_n;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sbase = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t a_hi = sbase;
          const uint32_t a_lo = sbase + p.a_tile_bytes;
          const uint32_t b_hi = sbase + p.nparts * p.a_tile_bytes;
          const uint32_t b_lo = b_hi + p.b_tile_bytes;
#pragma unroll
          for (int k = 0; k < SW_BYTES / 32; ++k) {  // 32 bytes of K per instruction
            const uint32_t koff = k * 32;
            const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
            if (p.nparts == 2) {
              // small cross terms first, the dominant hi*hi term last
              umma<KIND>(d_tmem, make_smem_desc(a_lo + koff), make_smem_desc(b_hi + koff), p.idesc, acc);
              umma<KIND>(d_tmem, make_smem_desc(a_hi + koff), make_smem_desc(b_lo + koff), p.idesc, 1u);
              umma<KIND>(d_tmem, make_smem_desc(a_hi + koff), make_smem_desc(b_hi + koff), p.idesc, 1u);
            } else {
              umma<KIND>(d_tmem, make_smem_desc(a_hi + koff), make_smem_desc(b_hi + koff), p.idesc, acc);
            }
          }
          tcgen05_commit(&empty_bar[stage]);  // frees this smem stage when the MMAs retire
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit(&tmem_full_bar[as]);  // accumulator complete
      }
    }
  } else {
    // ================= epilogue (warps 2..5) =================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const bool vec_ok = (p.c_cs == 1) && ((p.c_rs & 3) == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    uint32_t it = 0;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
    const uint32_t as = it % (uint32_t)p.acc_stages;
    const uint32_t aphase = (it / (uint32_t)p.acc_stages) & 1u;
    const long long m0 = (tile / tiles_n) * BLOCK_M;
    const long long n0 = (tile % tiles_n) * p.block_n;
    mbar_wait(&tmem_full_bar[as], aphase);
    tcgen05_fence_after();
    const long long row = m0 + q * 32 + lane;
    const uint32_t t_acc = tmem_base + as * (uint32_t)p.block_n + ((uint32_t)(q * 32) << 16);
    for (int c0 = 0; c0 < p.block_n; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(t_acc + (uint32_t)c0, r);
      if (c0 + 32 >= p.block_n) {
        // last chunk is in registers: hand the accumulator stage back to the MMA warp
        tcgen05_fence_before();
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[as]))
                     : "memory");
      }
      if (row < p.M) {
        const long long col0 = n0 + c0;
        float* crow = p.C + row * p.c_rs;
        if (vec_ok && col0 + 32 <= p.N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 v;
            v.x = p.alpha * __uint_as_float(r[j]);
            v.y = p.alpha * __uint_as_float(r[j + 1]);
            v.z = p.alpha * __uint_as_float(r[j + 2]);
            v.w = p.alpha * __uint_as_float(r[j + 3]);
            float4* dst = reinterpret_cast<float4*>(crow + col0 + j);
            if (p.beta != 0.0f) {
              const float4 o = *dst;
              v.x += p.beta * o.x; v.y += p.beta * o.y; v.z += p.beta * o.z; v.w += p.beta * o.w;
            }
            *dst = v;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const long long col = col0 + j;
            if (col < p.N) {
              float* dst = crow + col * p.c_cs;
              float v = p.alpha * __uint_as_float(r[j]);
              if (p.beta != 0.0f) v += p.beta * (*dst);
              *dst = v;
            }
          }
        }
      }
    }
    }  // tile loop
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(tmem_cols)
                 : "memory");
  }
}

// ------------------------------------------------------------------ operand packing
// out planes are [R, pitch] row-major (K-major): out[r*pitch + c] = f(in[r*s_r + c*s_c]).
// MODE 0: hi/lo tf32 split (two f32 planes), 1: f32 copy, 2: bf16.
template <int MODE>
__global__ void __launch_bounds__(256)
pack_kernel(const float* __restrict__ in, long long R, long long Kc, long long s_r, long long s_c,
            void* __restrict__ out0, void* __restrict__ out1, long long pitch) {
  __shared__ float tile[32][33];
  const long long r0 = (long long)blockIdx.y * 32, c0 = (long long)blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const bool r_fast = (s_r == 1) || (s_c != 1 && llabs(s_r) < llabs(s_c));
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    long long r, c;
    if (r_fast) { r = r0 + tx; c = c0 + ty + i; } else { r = r0 + ty + i; c = c0 + tx; }
    float v = 0.0f;
    if (r < R && c < Kc) v = in[r * s_r + c * s_c];
    if (r_fast) tile[tx][ty + i] = v; else tile[ty + i][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const long long r = r0 + ty + i, c = c0 + tx;
    if (r < R && c < Kc) {
      const float v = tile[ty + i][tx];
      if (MODE == 0) {
        uint32_t hi_bits, lo_bits;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi_bits) : "f"(v));
        const float hi = __uint_as_float(hi_bits);
        const float lo = v - hi;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo_bits) : "f"(lo));
        static_cast<float*>(out0)[r * pitch + c] = hi;
        static_cast<float*>(out1)[r * pitch + c] = __uint_as_float(lo_bits);
      } else if (MODE == 1) {
        static_cast<float*>(out0)[r * pitch + c] = v;
      } else {
        static_cast<__nv_bfloat16*>(out0)[r * pitch + c] = __float2bfloat16_rn(v);
      }
    }
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  }
  return fn;
}

int make_map(CUtensorMap* map, const void* base, bool bf16, long long rows, long long kc,
             long long pitch_elems, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(AB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const int es = bf16 ? 2 : 4;
  cuuint64_t dims[2] = {(cuuint64_t)kc, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch_elems * es};
  cuuint32_t box[2] = {(cuuint32_t)(SW_BYTES / es), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(AB_ERR_CUDA, "cuTensorMapEncodeTiled failed with code %d", (int)r);
  return AB_OK;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct OperandPlan {
  bool pack;          // needs the pack pre-pass
  long long pitch;    // elements per packed row
  size_t plane_bytes; // bytes of one packed plane
};

// logical operand [R, Kc] with element strides (s_r, s_c)
OperandPlan plan_operand(int precision, long long R, long long Kc, long long s_r, long long s_c,
                         const void* ptr) {
  OperandPlan o{};
  const bool direct_ok = (precision == 1) && (s_c == 1 || Kc == 1) && (s_r % 4 == 0) && s_r >= Kc &&
                         (reinterpret_cast<uintptr_t>(ptr) % 16 == 0);
  o.pack = !direct_ok;
  const int es = precision == 2 ? 2 : 4;
  o.pitch = (long long)align_up((size_t)Kc, 16 / es);
  o.plane_bytes = o.pack ? align_up((size_t)R * o.pitch * es, 1024) : 0;
  return o;
}

bool eligible(long long M, long long N, long long K) {
  return M >= 64 && N >= 64 && K >= 32 && (double)M * N * K >= (double)(1 << 21) &&
         M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31);
}

}  // namespace

size_t gemm_tcgen05_workspace(int precision, long long M, long long N, long long K, long long a_rs,
                              long long a_cs, long long b_rs, long long b_cs) {
  if (!eligible(M, N, K)) return 0;
  // pointer alignment is unknown here: assume packing is needed
  const int parts = precision == 0 ? 2 : 1;
  OperandPlan a = plan_operand(precision, M, K, a_rs, a_cs, reinterpret_cast<const void*>(1));
  OperandPlan b = plan_operand(precision, N, K, b_cs, b_rs, reinterpret_cast<const void*>(1));
  return parts * (a.plane_bytes + b.plane_bytes) + 2048;
}

int gemm_tcgen05_f32(int precision, long long M, long long N, long long K, float alpha,
                     const float* A, long long a_rs, long long a_cs, const float* B, long long b_rs,
                     long long b_cs, float beta, float* C, long long c_rs, long long c_cs,
                     void* workspace, size_t workspace_bytes, cudaStream_t st, bool* handled) {
  *handled = false;
  if (precision < 0 || precision > 2) return fail(AB_ERR_INVALID, "bad gemm precision %d", precision);
  if (!eligible(M, N, K)) return AB_OK;
  *handled = true;
  const int parts = precision == 0 ? 2 : 1;
  const bool bf16 = precision == 2;
  OperandPlan pa = plan_operand(precision, M, K, a_rs, a_cs, A);
  OperandPlan pb = plan_operand(precision, N, K, b_cs, b_rs, B);
  const size_t need = parts * (pa.plane_bytes + pb.plane_bytes) + 2048;
  if ((pa.pack || pb.pack) && (!workspace || workspace_bytes < need))
    return fail(AB_ERR_INVALID, "Gemm workspace too small: need %zu bytes, have %zu", need, workspace_bytes);
  uint8_t* ws = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(workspace), 1024));
  const void* a_plane[2] = {A, nullptr};
  const void* b_plane[2] = {B, nullptr};
  long long a_pitch = a_rs, b_pitch = b_cs;
  auto run_pack = [&](const float* src, long long R, long long s_r, long long s_c, OperandPlan& pl,
                      const void** planes, long long* pitch) -> int {
    void* o0 = ws;
    void* o1 = parts == 2 ? ws + pl.plane_bytes : nullptr;
    ws += parts * pl.plane_bytes;
    dim3 grid((unsigned)((K + 31) / 32), (unsigned)((R + 31) / 32));
    if (grid.y > 65535) return fail(AB_ERR_UNSUPPORTED, "gemm operand with more than 2M rows");
    if (precision == 0) pack_kernel<0><<<grid, 256, 0, st>>>(src, R, K, s_r, s_c, o0, o1, pl.pitch);
    else if (precision == 1) pack_kernel<1><<<grid, 256, 0, st>>>(src, R, K, s_r, s_c, o0, o1, pl.pitch);
    else pack_kernel<2><<<grid, 256, 0, st>>>(src, R, K, s_r, s_c, o0, o1, pl.pitch);
    g_launches++;
    AB_CUDA(cudaGetLastError());
    planes[0] = o0;
    planes[1] = o1;
    *pitch = pl.pitch;
    return AB_OK;
  };
  int rc;
  if (pa.pack && (rc = run_pack(A, M, a_rs, a_cs, pa, a_plane, &a_pitch))) return rc;
  if (pb.pack && (rc = run_pack(B, N, b_cs, b_rs, pb, b_plane, &b_pitch))) return rc;

  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.alpha = alpha; p.beta = beta;
  p.C = C; p.c_rs = c_rs; p.c_cs = c_cs;
  p.block_n = N >= 256 ? 256 : (N >= 128 ? 128 : 64);
  p.nparts = parts;
  p.k_elems_per_row = bf16 ? 64 : 32;
  p.a_tile_bytes = BLOCK_M * SW_BYTES;
  p.b_tile_bytes = p.block_n * SW_BYTES;
  const int stage_bytes = parts * (p.a_tile_bytes + p.b_tile_bytes);
  p.stages = std::max(2, std::min(8, (kMaxSmem - 1024) / stage_bytes));
  p.acc_stages = 2;  // 2 x block_n <= 512 TMEM columns
  // cute UMMA::InstrDescriptor: c_format F32=1 @[4,6), a/b format @[7,10)/[10,13)
  // (TF32=2, BF16=1), K-major both (@15, @16 = 0), N>>3 @[17,23), M>>4 @[24,29)
  const uint32_t fmt = bf16 ? 1u : 2u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.block_n >> 3) << 17) |
            ((uint32_t)(BLOCK_M >> 4) << 24);

  CUtensorMap ma0, ma1, mb0, mb1;
  if ((rc = make_map(&ma0, a_plane[0], bf16, M, K, a_pitch, BLOCK_M))) return rc;
  if ((rc = make_map(&mb0, b_plane[0], bf16, N, K, b_pitch, p.block_n))) return rc;
  ma1 = ma0; mb1 = mb0;
  if (parts == 2) {
    if ((rc = make_map(&ma1, a_plane[1], bf16, M, K, a_pitch, BLOCK_M))) return rc;
    if ((rc = make_map(&mb1, b_plane[1], bf16, N, K, b_pitch, p.block_n))) return rc;
  }
  const size_t smem = (size_t)p.stages * stage_bytes + 1024;
  const long long num_tiles = ((N + p.block_n - 1) / p.block_n) * ((M + BLOCK_M - 1) / BLOCK_M);
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  dim3 grid((unsigned)std::min<long long>(num_tiles, sms));  // persistent: one CTA per SM
  if (bf16) {
    static bool attr1 = false;
    if (!attr1) {
      AB_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
      attr1 = true;
    }
    gemm_tcgen05_kernel<1><<<grid, kThreads, smem, st>>>(ma0, ma1, mb0, mb1, p);
  } else {
    static bool attr0 = false;
    if (!attr0) {
      AB_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
      attr0 = true;
    }
    gemm_tcgen05_kernel<0><<<grid, kThreads, smem, st>>>(ma0, ma1, mb0, mb1, p);
  }
  g_launches++;
  AB_CUDA(cudaGetLastError());
  return AB_OK;
}

}  // namespace ab
