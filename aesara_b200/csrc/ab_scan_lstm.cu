// ab_scan_lstm.cu — Scan fast path: an LSTM-cell recurrence as ONE persistent kernel.
//
// Reference: aesara/scan/op.py:637 (Scan), loop :1799-2103 — for every step the
// reference calls an inner compiled function (Gemm + four column-slice Subtensors + two
// fused Elemwise nodes) from a Python/Cython loop.  When
// runtime/scan.py recognises that inner graph
//     pre = x_t + h_{t-1} @ U                      (Gemm, alpha = beta = 1)
//     c_t = sigmoid(pre_f) * c_{t-1} + sigmoid(pre_i) * tanh(pre_g)
//     h_t = sigmoid(pre_o) * tanh(c_t)              (gate order i, f, o, g)
// the whole T-step loop runs here without returning to the host:
//
//   * the grid is one CTA per SM (cooperative launch, all CTAs co-resident); every step
//     the [B, 4H] pre-activation is produced as 128 x 256 tcgen05 tiles (3xTF32,
//     fp32-faithful like ab_gemm) whose 256 columns are the four gates of 64 hidden units
//     (U is packed once with its columns gate-interleaved);
//   * the LSTM cell is the tile's epilogue: accumulators come out of TMEM, x_t and
//     c_{t-1} are read once, c_t / h_t are written to the Scan's circular output buffers
//     and h_t is ALSO written as the hi/lo TF32 planes the next step's TMA loads — the
//     pre-activations never touch HBM and there is no per-step pack or copy kernel;
//     K is accumulated in 128-element segments (fresh TMEM accumulator each, summed in
//     FP32 registers) because the tensor core's own accumulate truncates;
//   * there is no barrier between steps.  Batch rows are independent, so a tile of step
//     t+1 only needs the h_t rows of ITS row block: every finished tile is counted on its
//     row block (release: bar.sync of the epilogue warps, __threadfence, atomicAdd) and the
//     TMA producer of a step-t+1 tile waits until all tiles of that row block are counted
//     (acquire load, then fence.proxy.async before the loads).  The h planes are double
//     buffered by step parity; a step-t+1 tile overwrites rows that only step-t tiles of
//     the same row block read, and those are complete by then.
//
// State lives in L2 (h planes 2 x 2 x B x H x 4 B, c in the output ring): 8192 x 1024
// state rows do not fit the register file / shared memory of 148 SMs at M = 128 tiles.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdio>
#include <vector>

#include "ab_common.h"
#include "ab_tcgen05.cuh"

namespace ab {

namespace {

using namespace ab::tc;

constexpr int UNITS = 64;            // hidden units per tile
constexpr int TILE_N = 4 * UNITS;    // 256 accumulator columns = 4 gates x 64 units
constexpr int KB = 32;               // K elements (fp32/tf32) per 128-byte smem row
constexpr int SEG_KB = 8;            // k-blocks per tensor-core accumulation segment (256 K):
                                     // two TMEM stages of MMA work cover one cell epilogue

struct LstmParams {
  long long T, B, H;
  const float* x;            // [T, B, 4H]
  long long x_ts, x_rs;      // element strides of x: step, row (columns contiguous)
  float* hbuf;               // Scan output ring of h: [S_h, B, H] contiguous rows
  float* cbuf;               // Scan output ring of c: [S_c, B, H]
  long long sh, sc;          // ring lengths (store_steps)
  long long pos_h, pos_c;    // ring position written at step 0
  float* hplane[2][2];       // [set][hi/lo] K-major planes [B, H] of h for the tensor cores
  unsigned int* row_done;    // per row block: tiles completed so far, all steps (zero-initialised)
  long long* trace;          // AB_LSTM_TRACE: per-CTA wait/busy cycle counters (or null)
  int stages;
  int a_tile_bytes, b_tile_bytes;
  uint32_t idesc;
};

__device__ __forceinline__ float sigmoidf_ref(float v) { return 1.0f / (1.0f + expf(-v)); }

// acc (+)= 32 accumulator columns of this thread's TMEM lane (round-to-nearest adds)
__device__ __forceinline__ void fold32(float (&acc)[32], uint32_t taddr) {
  uint32_t r[32];
  tmem_ld_32x32b_x32(taddr, r);
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(r[j]);
}

// CTAS = 1: one CTA per 128 x 256 tile.  CTAS = 2: a cluster of two CTAs (cta_group::2)
// per 256 x 256 tile — each CTA stages its 128 batch rows of h and HALF of the U tile, so
// a pipeline stage is 64 KB instead of 96 KB (three stages instead of two in the 200 KB
// of shared memory, and a third less L2->SM traffic); the 1-CTA kernel spent more than
// half of its time waiting on its two-stage ring (profiles/r01_lstm_scan_v1.txt).
template <int CTAS>
__global__ void __launch_bounds__(kThreads, 1)
lstm_scan_kernel(const __grid_constant__ CUtensorMap map_h00, const __grid_constant__ CUtensorMap map_h01,
                 const __grid_constant__ CUtensorMap map_h10, const __grid_constant__ CUtensorMap map_h11,
                 const __grid_constant__ CUtensorMap map_u0, const __grid_constant__ CUtensorMap map_u1,
                 const __grid_constant__ LstmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  __shared__ __align__(8) uint64_t full_bar[8];
  __shared__ __align__(8) uint64_t empty_bar[8];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  constexpr bool TWO = CTAS == 2;
  constexpr int TILE_M = CTAS * BLOCK_M;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = TWO ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const long long group_id = blockIdx.x / CTAS, n_groups = gridDim.x / CTAS;
  const int stage_bytes = 2 * (p.a_tile_bytes + p.b_tile_bytes);  // hi + lo of A and B
  const int num_k_blocks = (int)((p.H + KB - 1) / KB);
  const long long tiles_n = p.H / UNITS;
  const long long num_tiles = ((p.B + TILE_M - 1) / TILE_M) * tiles_n;
  const uint32_t tmem_cols = 2 * TILE_N;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], CTAS * (kThreads - 64));  // every epilogue thread of the group
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if (TWO) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32(&tmem_base_slot)),
                   "r"(tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32(&tmem_base_slot)),
                   "r"(tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (TWO) cluster_sync_all();  // both CTAs' barriers exist before any remote signal
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  long long tr_a = 0, tr_b = 0, tr_c = 0, tr_d = 0;  // AB_LSTM_TRACE counters (see the end)
  const long long kernel_t0 = clock64();
  // pipeline state persists across tiles and steps
  int stage = 0;
  uint32_t phase = 0;
  uint32_t it = 0;

  for (long long t = 0; t < p.T; ++t) {
    const int set = (int)(t & 1);  // h planes read this step; the other set is written
    if (warp == 0) {
      // ================= TMA producer (one per CTA) =================
      if (lane == 0) {
        const CUtensorMap* mh0 = set ? &map_h10 : &map_h00;
        const CUtensorMap* mh1 = set ? &map_h11 : &map_h01;
        for (long long tile = group_id; tile < num_tiles; tile += n_groups) {
          const int m0 = (int)((tile / tiles_n) * TILE_M) + (int)rank * BLOCK_M;
          const int n0 = (int)((tile % tiles_n) * TILE_N) + (int)rank * (TILE_N / 2);
          if (t > 0) {
            // rows of h_{t-1} for this row block exist once all of its step t-1 tiles are
            // counted (tiles_n tiles x CTAS signalling CTAs per step)
            const unsigned int target = (unsigned int)(t * tiles_n * CTAS);
            const unsigned int* flag = p.row_done + tile / tiles_n;
            const long long w0 = clock64();
            unsigned int seen;
            do {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
              if (seen < target) __nanosleep(32);
            } while (seen < target);
            asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy writes -> TMA reads
            tr_c += clock64() - w0;
          }
          for (int kb = 0; kb < num_k_blocks; ++kb) {
            const long long c0 = clock64();
            mbar_wait(&empty_bar[stage], phase ^ 1);
            tr_a += clock64() - c0;
            uint8_t* sbase = smem + (size_t)stage * stage_bytes;
            if (leader) mbar_expect_tx(&full_bar[stage], (uint32_t)(CTAS * stage_bytes));
            const int kc = kb * KB;
            uint8_t* a_hi = sbase;
            uint8_t* a_lo = sbase + p.a_tile_bytes;
            uint8_t* b_hi = sbase + 2 * p.a_tile_bytes;
            uint8_t* b_lo = b_hi + p.b_tile_bytes;
            if (TWO) {
              tma_load_2d_2sm(a_hi, mh0, &full_bar[stage], kc, m0);
              tma_load_2d_2sm(a_lo, mh1, &full_bar[stage], kc, m0);
              tma_load_2d_2sm(b_hi, &map_u0, &full_bar[stage], kc, n0);
              tma_load_2d_2sm(b_lo, &map_u1, &full_bar[stage], kc, n0);
            } else {
              tma_load_2d(a_hi, mh0, &full_bar[stage], kc, m0);
              tma_load_2d(a_lo, mh1, &full_bar[stage], kc, m0);
              tma_load_2d(b_hi, &map_u0, &full_bar[stage], kc, n0);
              tma_load_2d(b_lo, &map_u1, &full_bar[stage], kc, n0);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == 1) {
      // ================= MMA issuer (the leader CTA of a pair) =================
      if (leader && lane == 0) {
        for (long long tile = group_id; tile < num_tiles; tile += n_groups) {
          // K segments: a fresh TMEM accumulator every SEG_KB k-blocks, folded into FP32
          // registers by the epilogue (the tensor core's accumulate truncates; see
          // ab_gemm_tcgen05.cu "segments")
          for (int kb0 = 0; kb0 < num_k_blocks; kb0 += SEG_KB, ++it) {
            const int kb1 = min(kb0 + SEG_KB, num_k_blocks);
            const uint32_t as = it & 1u, aphase = (it >> 1) & 1u;
            long long c0 = clock64();
            mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
            tr_a += clock64() - c0;
            tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + as * TILE_N;
            for (int kb = kb0; kb < kb1; ++kb) {
              c0 = clock64();
              mbar_wait(&full_bar[stage], phase);
              tr_b += clock64() - c0;
              tcgen05_fence_after();
              const uint32_t sbase = smem_u32(smem + (size_t)stage * stage_bytes);
              const uint32_t a_hi = sbase, a_lo = sbase + p.a_tile_bytes;
              const uint32_t b_hi = sbase + 2 * p.a_tile_bytes, b_lo = b_hi + p.b_tile_bytes;
#pragma unroll
              for (int k = 0; k < SW_BYTES / 32; ++k) {
                const uint32_t ko = k * 32;
                const uint32_t acc = (kb > kb0 || k > 0) ? 1u : 0u;
                if (TWO) {
                  umma_2sm<0>(d_tmem, make_smem_desc(a_lo + ko, 16), make_smem_desc(b_hi + ko, 16), p.idesc, acc);
                  umma_2sm<0>(d_tmem, make_smem_desc(a_hi + ko, 16), make_smem_desc(b_lo + ko, 16), p.idesc, 1u);
                  umma_2sm<0>(d_tmem, make_smem_desc(a_hi + ko, 16), make_smem_desc(b_hi + ko, 16), p.idesc, 1u);
                } else {
                  umma<0>(d_tmem, make_smem_desc(a_lo + ko, 16), make_smem_desc(b_hi + ko, 16), p.idesc, acc);
                  umma<0>(d_tmem, make_smem_desc(a_hi + ko, 16), make_smem_desc(b_lo + ko, 16), p.idesc, 1u);
                  umma<0>(d_tmem, make_smem_desc(a_hi + ko, 16), make_smem_desc(b_hi + ko, 16), p.idesc, 1u);
                }
              }
              if (TWO) tcgen05_commit_2sm(&empty_bar[stage]);
              else tcgen05_commit(&empty_bar[stage]);
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            if (TWO) tcgen05_commit_2sm(&tmem_full_bar[as]);
            else tcgen05_commit(&tmem_full_bar[as]);
          }
        }
      }
    } else {
      // ================= epilogue = the LSTM cell (warps 2..9) =================
      // two warps per TMEM lane quarter; warp group s owns hidden units [32 s, 32 s + 32)
      // of the tile, i.e. 4 gates x 32 accumulator columns per thread
      const int q = warp & 3;
      const int s = (warp - 2) >> 2;
      const long long rh = (p.pos_h + t) % p.sh;                   // ring rows written now
      const long long rc = (p.pos_c + t) % p.sc;
      const long long rc_prev = (p.pos_c + t - 1 + p.sc) % p.sc;   // c_{t-1}
      float* h_out = p.hbuf + rh * p.B * p.H;
      float* c_out = p.cbuf + rc * p.B * p.H;
      const float* c_in = p.cbuf + rc_prev * p.B * p.H;
      float* hp_hi = p.hplane[set ^ 1][0];
      float* hp_lo = p.hplane[set ^ 1][1];
      const float* xt = p.x + t * p.x_ts;
      for (long long tile = group_id; tile < num_tiles; tile += n_groups) {
        const long long m0 = (tile / tiles_n) * TILE_M + (long long)rank * BLOCK_M;
        const long long u0 = (tile % tiles_n) * UNITS;
        const long long row = m0 + q * 32 + lane;
        // The accumulator registers start out as this row's x_t gate pre-activations: all
        // 32 loads are in flight while the tensor core produces the first K segment, and
        // the Gemm's "+ 1 * x_t" (blas.py:984-1017) costs no registers or latency later.
        float ai[32], af[32], ao[32], ag[32];
        const long long uc = u0 + s * 32;
        const long long so = row * p.H + uc;
        if (row < p.B) {
          const float* xr = xt + row * p.x_rs + uc;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 xi = __ldcs(reinterpret_cast<const float4*>(xr + 0 * p.H + j));
            const float4 xf = __ldcs(reinterpret_cast<const float4*>(xr + 1 * p.H + j));
            const float4 xo = __ldcs(reinterpret_cast<const float4*>(xr + 2 * p.H + j));
            const float4 xg = __ldcs(reinterpret_cast<const float4*>(xr + 3 * p.H + j));
            ai[j] = xi.x; ai[j + 1] = xi.y; ai[j + 2] = xi.z; ai[j + 3] = xi.w;
            af[j] = xf.x; af[j + 1] = xf.y; af[j + 2] = xf.z; af[j + 3] = xf.w;
            ao[j] = xo.x; ao[j + 1] = xo.y; ao[j + 2] = xo.z; ao[j + 3] = xo.w;
            ag[j] = xg.x; ag[j + 1] = xg.y; ag[j + 2] = xg.z; ag[j + 3] = xg.w;
          }
          // c_{t-1} of this row/unit range is needed after the last segment: pull it into L2
          asm volatile("prefetch.global.L2 [%0];" ::"l"(c_in + so));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) ai[j] = af[j] = ao[j] = ag[j] = 0.0f;
        }
        for (int kb0 = 0; kb0 < num_k_blocks; kb0 += SEG_KB, ++it) {
          const uint32_t as = it & 1u, aphase = (it >> 1) & 1u;
          const long long c0 = clock64();
          mbar_wait(&tmem_full_bar[as], aphase);
          tr_a += clock64() - c0;
          tcgen05_fence_after();
          const uint32_t t_acc = tmem_base + as * TILE_N + ((uint32_t)(q * 32) << 16) + s * 32;
          fold32(ai, t_acc + 0 * UNITS);
          fold32(af, t_acc + 1 * UNITS);
          fold32(ao, t_acc + 2 * UNITS);
          fold32(ag, t_acc + 3 * UNITS);
          tcgen05_fence_before();
          if (TWO) {
            asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(
                             smem_u32(&tmem_empty_bar[as]) & kPeerBitMask)
                         : "memory");
          } else {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[as]))
                         : "memory");
          }
        }
        const long long cell0 = clock64();
        if (row < p.B) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 cp = *reinterpret_cast<const float4*>(c_in + so + j);
            float cn[4], hn[4], hh[4], hl[4];
            const float cpv[4] = {cp.x, cp.y, cp.z, cp.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              // the two Elemwise composites of the inner graph on pre = x_t + h_{t-1} @ U
              cn[e] = sigmoidf_ref(af[j + e]) * cpv[e] + sigmoidf_ref(ai[j + e]) * tanhf(ag[j + e]);
              hn[e] = sigmoidf_ref(ao[j + e]) * tanhf(cn[e]);
              uint32_t hb, lb;
              asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(hn[e]));
              hh[e] = __uint_as_float(hb);
              asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(hn[e] - hh[e]));
              hl[e] = __uint_as_float(lb);
            }
            *reinterpret_cast<float4*>(c_out + so + j) = make_float4(cn[0], cn[1], cn[2], cn[3]);
            *reinterpret_cast<float4*>(h_out + so + j) = make_float4(hn[0], hn[1], hn[2], hn[3]);
            *reinterpret_cast<float4*>(hp_hi + so + j) = make_float4(hh[0], hh[1], hh[2], hh[3]);
            *reinterpret_cast<float4*>(hp_lo + so + j) = make_float4(hl[0], hl[1], hl[2], hl[3]);
          }
        }
        tr_b += clock64() - cell0;
        // h_t / c_t of this tile are written: count the tile on its row block so that the
        // producers of step t+1 may load those rows (all 256 epilogue threads' stores first)
        asm volatile("bar.sync 1, %0;" ::"r"(kThreads - 64) : "memory");
        if (warp == 2 && lane == 0) {
          __threadfence();
          atomicAdd(p.row_done + tile / tiles_n, 1u);
        }
      }
    }
  }
  if (p.trace && lane == 0 && (warp == 0 || warp == 1 || warp == 2)) {
    // [cta][role 0..2][4]: producer {empty-wait, -, row-block wait, -},
    // MMA {tmem-empty-wait, smem-full-wait}, epilogue warp 2 {tmem-full-wait, cell time}; [3] = total
    long long* o = p.trace + ((long long)blockIdx.x * 3 + warp) * 5;
    o[0] = tr_a; o[1] = tr_b; o[2] = tr_c; o[3] = tr_d; o[4] = clock64() - kernel_t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (TWO) cluster_sync_all();  // nobody signals the peer's barriers / reads its smem after this
  if (warp == 1) {
    tcgen05_fence_after();
    if (TWO) {
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                   "r"(tmem_cols)
                   : "memory");
    } else {
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                   "r"(tmem_cols)
                   : "memory");
    }
  }
}

// hi/lo TF32 planes of a [R, K] row-major-able matrix; rows optionally gate-interleaved:
// plane row n' = tile*256 + gate*64 + u  <-  source column gate*H + tile*64 + u of U[K, 4H]
__global__ void __launch_bounds__(256)
split_planes_kernel(const float* __restrict__ src, long long R, long long K, long long s_r,
                    long long s_k, int interleave_h, float* __restrict__ hi, float* __restrict__ lo) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= R * K) return;
  const long long r = i / K, k = i - r * K;
  long long rs = r;
  if (interleave_h > 0) {
    const long long tile = r / TILE_N, rem = r % TILE_N;
    const long long gate = rem / UNITS, u = rem % UNITS;
    rs = gate * interleave_h + tile * UNITS + u;
  }
  const float v = src[rs * s_r + k * s_k];
  uint32_t hb, lb;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
  const float h = __uint_as_float(hb);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(v - h));
  hi[i] = h;
  lo[i] = __uint_as_float(lb);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map_f32(CUtensorMap* map, const void* base, long long k, long long rows, int box_rows) {
  static EncodeTiledFn enc = nullptr;
  if (!enc) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return fail(AB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    enc = reinterpret_cast<EncodeTiledFn>(fp);
  }
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)k * 4};
  cuuint32_t box[2] = {(cuuint32_t)KB, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(AB_ERR_CUDA, "cuTensorMapEncodeTiled failed with code %d", (int)r);
  return AB_OK;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline size_t row_flag_bytes(long long b) { return align_up((size_t)((b + BLOCK_M - 1) / BLOCK_M) * 4, 256); }

// AB_LSTM_TRACE=1: wait for the kernel and print where each role's cycles went (debug aid)
void dump_trace(long long* dev, int ctas, cudaStream_t st) {
  std::vector<long long> h((size_t)ctas * 15);
  cudaStreamSynchronize(st);
  cudaMemcpy(h.data(), dev, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaFree(dev);
  const char* names[3][4] = {{"producer: smem-empty wait", "-", "row-block (h_t-1) wait", "-"},
                             {"mma: tmem-empty wait", "mma: smem-full wait", "-", "-"},
                             {"epilogue: tmem-full wait", "epilogue: cell+stores", "-", "-"}};
  double total = 0;
  for (int c = 0; c < ctas; ++c) total += (double)h[(size_t)c * 15 + 4];
  total /= ctas;
  fprintf(stderr, "[ab_lstm_scan trace] %d CTAs, %.0f cycles per CTA\n", ctas, total);
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 4; ++k) {
      if (names[r][k][0] == '-') continue;
      double sum = 0, mx = 0;
      int n = 0;
      for (int c = 0; c < ctas; ++c) {
        const double v = (double)h[((size_t)c * 3 + r) * 5 + k];
        if (r == 1 && h[((size_t)c * 3 + 1) * 5] == 0 && h[((size_t)c * 3 + 1) * 5 + 1] == 0) continue;  // non-leader CTA
        sum += v; mx = std::max(mx, v); ++n;
      }
      if (n) fprintf(stderr, "  %-28s avg %5.1f %%  max %5.1f %%\n", names[r][k], 100.0 * sum / n / total, 100.0 * mx / total);
    }
}

}  // namespace
}  // namespace ab

using namespace ab;

extern "C" int ab_lstm_scan_workspace_bytes(int64_t b, int64_t h, size_t* bytes) {
  if (!bytes) return fail(AB_ERR_INVALID, "null out pointer");
  const size_t hp = align_up((size_t)b * h * 4, 1024);
  const size_t up = align_up((size_t)4 * h * h * 4, 1024);
  *bytes = 4 * hp + 2 * up + 1024 + row_flag_bytes(b);
  return AB_OK;
}

extern "C" int ab_lstm_scan_supported(int64_t t, int64_t b, int64_t h) {
  return (t >= 1 && b >= 1 && h >= UNITS && h % UNITS == 0 && h % 4 == 0 && b < (1LL << 31) &&
          h < (1LL << 29) && t * (h / UNITS) * 2 < (1LL << 31)) ? 1 : 0;
}

// h0 [B,H] (strides h0_rs, 1...), c0 likewise are expected to be already written into the
// rings at row (pos - 1) by the caller (Scan's IncSubtensor{InplaceSet} initial-state
// placement); h_init points at that row of the h ring (contiguous [B,H]).
extern "C" int ab_lstm_scan(int64_t T, int64_t B, int64_t H, const void* x, int64_t x_ts,
                            int64_t x_rs, const void* U, int64_t u_rs, int64_t u_cs, void* hbuf,
                            int64_t sh, int64_t pos_h, void* cbuf, int64_t sc, int64_t pos_c,
                            void* workspace, size_t workspace_bytes, void* stream) {
  if (!ab_lstm_scan_supported(T, B, H)) return fail(AB_ERR_UNSUPPORTED, "LSTM scan shape not supported");
  if ((x_rs % 4) || (x_ts % 4) || (reinterpret_cast<uintptr_t>(x) % 16))
    return fail(AB_ERR_UNSUPPORTED, "x must be 16-byte aligned with 16-byte aligned rows");
  cudaStream_t st = as_stream(stream);
  size_t need = 0;
  ab_lstm_scan_workspace_bytes(B, H, &need);
  if (!workspace || workspace_bytes < need)
    return fail(AB_ERR_INVALID, "LSTM scan workspace too small: need %zu bytes, have %zu", need, workspace_bytes);
  uint8_t* ws = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(workspace), 1024));
  const size_t hp = align_up((size_t)B * H * 4, 1024);
  const size_t up = align_up((size_t)4 * H * H * 4, 1024);
  LstmParams p{};
  p.T = T; p.B = B; p.H = H;
  p.x = static_cast<const float*>(x); p.x_ts = x_ts; p.x_rs = x_rs;
  p.hbuf = static_cast<float*>(hbuf); p.cbuf = static_cast<float*>(cbuf);
  p.sh = sh; p.sc = sc; p.pos_h = pos_h; p.pos_c = pos_c;
  for (int s = 0; s < 2; ++s)
    for (int k = 0; k < 2; ++k) p.hplane[s][k] = reinterpret_cast<float*>(ws + (size_t)(2 * s + k) * hp);
  float* u_hi = reinterpret_cast<float*>(ws + 4 * hp);
  float* u_lo = reinterpret_cast<float*>(ws + 4 * hp + up);
  p.row_done = reinterpret_cast<unsigned int*>(ws + 4 * hp + 2 * up);
  AB_CUDA(cudaMemsetAsync(p.row_done, 0, row_flag_bytes(B), st));
  static const bool want_trace = getenv("AB_LSTM_TRACE") != nullptr;
  long long* trace_dev = nullptr;
  if (want_trace) {
    AB_CUDA(cudaMalloc(&trace_dev, sizeof(long long) * 15 * 1024));
    AB_CUDA(cudaMemsetAsync(trace_dev, 0, sizeof(long long) * 15 * 1024, st));
    p.trace = trace_dev;
  }
  // planes of h_{-1}: the ring row just before pos_h
  const float* h_init = p.hbuf + ((pos_h - 1 + sh) % sh) * B * H;
  {
    const long long n = B * H;
    split_planes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h_init, B, H, H, 1, 0,
                                                                    p.hplane[0][0], p.hplane[0][1]);
    g_launches++;
    // U[K=H, N=4H] -> planes [4H (gate-interleaved), K=H]: element (n, k) = U[k*u_rs + n*u_cs]
    const long long nu = 4 * H * H;
    split_planes_kernel<<<(unsigned)((nu + 255) / 256), 256, 0, st>>>(
        static_cast<const float*>(U), 4 * H, H, u_cs, u_rs, (int)H, u_hi, u_lo);
    g_launches++;
    AB_CUDA(cudaGetLastError());
  }
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  static const bool allow_2cta = getenv("AB_LSTM_1CTA") == nullptr;
  bool two_cta = allow_2cta && B >= 2 * BLOCK_M && sms % 2 == 0;
  CUtensorMap mh[2][2], mu[2];
  int rc;
  for (int s = 0; s < 2; ++s)
    for (int k = 0; k < 2; ++k)
      if ((rc = make_map_f32(&mh[s][k], p.hplane[s][k], H, B, BLOCK_M))) return rc;
  void* args[] = {&mh[0][0], &mh[0][1], &mh[1][0], &mh[1][1], &mu[0], &mu[1], &p};
  p.a_tile_bytes = BLOCK_M * SW_BYTES;
  static bool attr1 = false, attr2 = false;

  if (two_cta) {
    // a cluster of two CTAs per 256 x 256 tile; every cluster must be co-resident for the
    // device-wide barrier: cooperative launch, grid <= what the occupancy query admits
    p.b_tile_bytes = (TILE_N / 2) * SW_BYTES;
    const int stage_bytes = 2 * (p.a_tile_bytes + p.b_tile_bytes);
    p.stages = std::max(2, std::min(8, (kMaxSmem - 1024) / stage_bytes));
    p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TILE_N >> 3) << 17) |
              ((uint32_t)((2 * BLOCK_M) >> 4) << 24);
    if ((rc = make_map_f32(&mu[0], u_hi, H, 4 * H, TILE_N / 2))) return rc;
    if ((rc = make_map_f32(&mu[1], u_lo, H, 4 * H, TILE_N / 2))) return rc;
    if (!attr2) {
      AB_CUDA(cudaFuncSetAttribute(lstm_scan_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
      attr2 = true;
    }
    const size_t smem = (size_t)p.stages * stage_bytes + 1024;
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeCooperative;
    attr[1].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    cfg.gridDim = dim3((unsigned)sms);
    int max_clusters = 0;
    cudaError_t qe = cudaOccupancyMaxActiveClusters(&max_clusters, lstm_scan_kernel<2>, &cfg);
    const long long tiles2 = ((B + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * (H / UNITS);
    if (qe == cudaSuccess && max_clusters >= 1) {
      const long long clusters = std::max<long long>(1, std::min<long long>(tiles2, max_clusters));
      cfg.gridDim = dim3((unsigned)(2 * clusters));
      cudaError_t le = cudaLaunchKernelExC(&cfg, (const void*)lstm_scan_kernel<2>, args);
      if (le == cudaSuccess) {
        g_launches++;
        if (trace_dev) dump_trace(trace_dev, (int)(2 * clusters), st);
        return AB_OK;
      }
    }
    cudaGetLastError();  // cluster + cooperative launch not available here: one CTA per tile
    two_cta = false;
  }
  p.b_tile_bytes = TILE_N * SW_BYTES;
  const int stage_bytes = 2 * (p.a_tile_bytes + p.b_tile_bytes);
  p.stages = std::max(2, std::min(8, (kMaxSmem - 1024) / stage_bytes));
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TILE_N >> 3) << 17) |
            ((uint32_t)(BLOCK_M >> 4) << 24);
  if ((rc = make_map_f32(&mu[0], u_hi, H, 4 * H, TILE_N))) return rc;
  if ((rc = make_map_f32(&mu[1], u_lo, H, 4 * H, TILE_N))) return rc;
  if (!attr1) {
    AB_CUDA(cudaFuncSetAttribute(lstm_scan_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    attr1 = true;
  }
  const size_t smem = (size_t)p.stages * stage_bytes + 1024;
  int per_sm = 0;
  AB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lstm_scan_kernel<1>, kThreads, smem));
  if (per_sm < 1) return fail(AB_ERR_CUDA, "lstm_scan_kernel does not fit on an SM");
  const long long num_tiles = ((B + BLOCK_M - 1) / BLOCK_M) * (H / UNITS);
  // every CTA must be resident for the device-wide barrier: cooperative launch, <= 1 per SM
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(num_tiles, sms));
  AB_CUDA(cudaLaunchCooperativeKernel((const void*)lstm_scan_kernel<1>, dim3(grid), dim3(kThreads), args,
                                      smem, st));
  g_launches++;
  if (trace_dev) dump_trace(trace_dev, (int)grid, st);
  return AB_OK;
}
