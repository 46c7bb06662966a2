// ab_rowops.cu — Softmax / LogSoftmax / SoftmaxGrad / MaxAndArgmax (SURVEY.md §8f N1:
// the first ops a classifier MLP needs after the hot path).
//
// Reference: aesara/tensor/special.py:239 (Softmax; C loop :440-478: row max,
// exp(x - max), scale by 1/sum), :508 (LogSoftmax; :715-740), :13 (SoftmaxGrad.perform
// :38-43: dy*sm - sum(dy*sm)*sm) and aesara/tensor/math.py:126 (MaxAndArgmax.perform
// :164-186: first index of the maximum, NaN wins like np.argmax).
//
// The operand is viewed as [outer, R, inner] (C-contiguous) with the op applied along R:
//   inner == 1  one warp (R <= 1024) or one CTA per row: three passes over the row in
//               registers/L1 (max, sum, write), warp-shuffle + smem trees;
//   inner  > 1  one thread per (outer, inner) column, coalesced across `inner`.
// HBM-bound: 1 read + 1 write of the tensor (the row is re-read from L1/L2).
// Sums of exponentials accumulate in double for float32 rows (the reference's float
// running sum is itself only ~R*eps accurate).
#include <cmath>

#include "ab_common.h"

using namespace ab;

namespace {

constexpr int kThreads = 256;

template <typename T> struct Acc { typedef double type; };

template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    T o = __shfl_xor_sync(0xffffffffu, v, m);
    v = (o > v || (o != o)) ? o : v;  // NaN propagates
  }
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

// block-wide helpers (blockDim.x == kThreads)
template <typename T>
__device__ T block_max(T v, T* sm) {
  v = warp_max(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sm[w] = v;
  __syncthreads();
  T r = sm[0];
  for (int i = 1; i < kThreads / 32; ++i) r = (sm[i] > r || (sm[i] != sm[i])) ? sm[i] : r;
  return r;
}
__device__ double block_sum(double v, double* sm) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sm[w] = v;
  __syncthreads();
  double r = 0;
  for (int i = 0; i < kThreads / 32; ++i) r += sm[i];
  return r;
}

// MODE 0 softmax, 1 log-softmax, 2 softmax-grad (in = dy, in2 = sm)
template <typename T, int MODE, bool WARP>
__global__ void __launch_bounds__(kThreads)
rows_kernel(long long outer, long long R, const T* __restrict__ in, const T* __restrict__ in2,
            T* __restrict__ out) {
  __shared__ double smd[kThreads / 32];
  __shared__ T smt[kThreads / 32];
  long long row;
  int tid, nthr;
  if (WARP) {
    row = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    tid = threadIdx.x & 31;
    nthr = 32;
    if (row >= outer) return;
  } else {
    row = blockIdx.x;
    tid = threadIdx.x;
    nthr = kThreads;
  }
  const T* x = in + row * R;
  T* y = out + row * R;
  if (MODE == 2) {
    const T* s = in2 + row * R;
    double acc = 0;
    for (long long i = tid; i < R; i += nthr) acc += (double)x[i] * (double)s[i];
    acc = WARP ? warp_sum(acc) : block_sum(acc, smd);
    const T tot = (T)acc;
    for (long long i = tid; i < R; i += nthr) y[i] = x[i] * s[i] - tot * s[i];
    return;
  }
  T mx = x[0];
  for (long long i = tid; i < R; i += nthr) {
    const T v = x[i];
    mx = (v > mx || (v != v)) ? v : mx;
  }
  mx = WARP ? warp_max(mx) : block_max(mx, smt);
  double acc = 0;
  for (long long i = tid; i < R; i += nthr) acc += exp((double)(x[i] - mx));
  acc = WARP ? warp_sum(acc) : block_sum(acc, smd);
  if (MODE == 0) {
    const T inv = (T)(1.0 / acc);
    for (long long i = tid; i < R; i += nthr) y[i] = (T)exp((double)(x[i] - mx)) * inv;
  } else {
    const T lse = (T)log(acc);
    for (long long i = tid; i < R; i += nthr) y[i] = (T)(x[i] - mx) - lse;
  }
}

// inner > 1: thread per (o, j) column, elements R apart by `inner`
template <typename T, int MODE>
__global__ void __launch_bounds__(kThreads)
cols_kernel(long long outer, long long R, long long inner, const T* __restrict__ in,
            const T* __restrict__ in2, T* __restrict__ out) {
  const long long c = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (c >= outer * inner) return;
  const long long o = c / inner, j = c - o * inner;
  const T* x = in + o * R * inner + j;
  T* y = out + o * R * inner + j;
  if (MODE == 2) {
    const T* s = in2 + o * R * inner + j;
    double acc = 0;
    for (long long i = 0; i < R; ++i) acc += (double)x[i * inner] * (double)s[i * inner];
    const T tot = (T)acc;
    for (long long i = 0; i < R; ++i) y[i * inner] = x[i * inner] * s[i * inner] - tot * s[i * inner];
    return;
  }
  T mx = x[0];
  for (long long i = 1; i < R; ++i) {
    const T v = x[i * inner];
    mx = (v > mx || (v != v)) ? v : mx;
  }
  double acc = 0;
  for (long long i = 0; i < R; ++i) acc += exp((double)(x[i * inner] - mx));
  if (MODE == 0) {
    const T inv = (T)(1.0 / acc);
    for (long long i = 0; i < R; ++i) y[i * inner] = (T)exp((double)(x[i * inner] - mx)) * inv;
  } else {
    const T lse = (T)log(acc);
    for (long long i = 0; i < R; ++i) y[i * inner] = (T)(x[i * inner] - mx) - lse;
  }
}

// ---- max + argmax over the trailing (flattened) axis of a [outer, R] matrix ------------
template <typename T>
__device__ __forceinline__ bool better(T v, long long i, T bv, long long bi) {
  const bool vn = (v != v), bn = (bv != bv);
  if (vn != bn) return vn;          // NaN beats everything (np.argmax / np.max)
  if (!vn && v != bv) return v > bv;
  return i < bi;                    // first occurrence
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
argmax_rows_kernel(long long outer, long long R, const T* __restrict__ in, T* __restrict__ omax,
                   long long* __restrict__ oidx) {
  __shared__ T sv[kThreads];
  __shared__ long long si[kThreads];
  const long long row = blockIdx.x;
  const T* x = in + row * R;
  T bv = x[0];
  long long bi = 0;
  for (long long i = threadIdx.x; i < R; i += kThreads) {
    const T v = x[i];
    if (better(v, i, bv, bi)) { bv = v; bi = i; }
  }
  sv[threadIdx.x] = bv;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int s = kThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      if (better(sv[threadIdx.x + s], si[threadIdx.x + s], sv[threadIdx.x], si[threadIdx.x])) {
        sv[threadIdx.x] = sv[threadIdx.x + s];
        si[threadIdx.x] = si[threadIdx.x + s];
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (omax) omax[row] = sv[0];
    if (oidx) oidx[row] = si[0];
  }
}

template <typename T>
int softmax_impl(int mode, long long outer, long long R, long long inner, const void* in,
                 const void* in2, void* out, cudaStream_t st) {
  if (outer == 0 || R == 0 || inner == 0) return AB_OK;
  const T* x = static_cast<const T*>(in);
  const T* x2 = static_cast<const T*>(in2);
  T* y = static_cast<T*>(out);
  if (inner == 1) {
    if (R <= 1024) {
      const unsigned blocks = (unsigned)((outer + 7) / 8);
      if (mode == 0) rows_kernel<T, 0, true><<<blocks, kThreads, 0, st>>>(outer, R, x, x2, y);
      else if (mode == 1) rows_kernel<T, 1, true><<<blocks, kThreads, 0, st>>>(outer, R, x, x2, y);
      else rows_kernel<T, 2, true><<<blocks, kThreads, 0, st>>>(outer, R, x, x2, y);
    } else {
      if (outer > 2147483647LL) return fail(AB_ERR_UNSUPPORTED, "too many rows");
      const unsigned blocks = (unsigned)outer;
      if (mode == 0) rows_kernel<T, 0, false><<<blocks, kThreads, 0, st>>>(outer, R, x, x2, y);
      else if (mode == 1) rows_kernel<T, 1, false><<<blocks, kThreads, 0, st>>>(outer, R, x, x2, y);
      else rows_kernel<T, 2, false><<<blocks, kThreads, 0, st>>>(outer, R, x, x2, y);
    }
  } else {
    const long long cols = outer * inner;
    const unsigned blocks = (unsigned)((cols + kThreads - 1) / kThreads);
    if (mode == 0) cols_kernel<T, 0><<<blocks, kThreads, 0, st>>>(outer, R, inner, x, x2, y);
    else if (mode == 1) cols_kernel<T, 1><<<blocks, kThreads, 0, st>>>(outer, R, inner, x, x2, y);
    else cols_kernel<T, 2><<<blocks, kThreads, 0, st>>>(outer, R, inner, x, x2, y);
  }
  g_launches++;
  AB_CUDA(cudaGetLastError());
  return AB_OK;
}

template <typename T>
int argmax_impl(long long outer, long long R, const void* in, void* omax, void* oidx,
                cudaStream_t st) {
  if (outer == 0) return AB_OK;
  if (R == 0) return fail(AB_ERR_SHAPE, "attempt to get argmax of an empty sequence");
  if (outer > 2147483647LL) return fail(AB_ERR_UNSUPPORTED, "too many rows");
  argmax_rows_kernel<T><<<(unsigned)outer, kThreads, 0, st>>>(
      outer, R, static_cast<const T*>(in), static_cast<T*>(omax), static_cast<long long*>(oidx));
  g_launches++;
  AB_CUDA(cudaGetLastError());
  return AB_OK;
}

}  // namespace

extern "C" int ab_softmax(int dtype, int mode, int64_t outer, int64_t r, int64_t inner,
                          const void* in, const void* in2, void* out, void* stream) {
  if (mode < 0 || mode > 2) return fail(AB_ERR_INVALID, "bad softmax mode %d", mode);
  if (mode == 2 && !in2) return fail(AB_ERR_INVALID, "SoftmaxGrad needs two inputs");
  if (dtype == AB_F32) return softmax_impl<float>(mode, outer, r, inner, in, in2, out, as_stream(stream));
  if (dtype == AB_F64) return softmax_impl<double>(mode, outer, r, inner, in, in2, out, as_stream(stream));
  return fail(AB_ERR_UNSUPPORTED, "Softmax family supports float32/float64 only");
}

extern "C" int ab_max_and_argmax(int dtype, int64_t outer, int64_t r, const void* in, void* out_max,
                                 void* out_argmax, void* stream) {
  cudaStream_t st = as_stream(stream);
  switch (dtype) {
    case AB_F32: return argmax_impl<float>(outer, r, in, out_max, out_argmax, st);
    case AB_F64: return argmax_impl<double>(outer, r, in, out_max, out_argmax, st);
    case AB_I8: return argmax_impl<signed char>(outer, r, in, out_max, out_argmax, st);
    case AB_I16: return argmax_impl<short>(outer, r, in, out_max, out_argmax, st);
    case AB_I32: return argmax_impl<int>(outer, r, in, out_max, out_argmax, st);
    case AB_I64: return argmax_impl<long long>(outer, r, in, out_max, out_argmax, st);
    case AB_U8:
    case AB_BOOL: return argmax_impl<unsigned char>(outer, r, in, out_max, out_argmax, st);
    case AB_U16: return argmax_impl<unsigned short>(outer, r, in, out_max, out_argmax, st);
    case AB_U32: return argmax_impl<unsigned int>(outer, r, in, out_max, out_argmax, st);
    case AB_U64: return argmax_impl<unsigned long long>(outer, r, in, out_max, out_argmax, st);
    default: return fail(AB_ERR_UNSUPPORTED, "MaxAndArgmax: unsupported dtype code %d", dtype);
  }
}
