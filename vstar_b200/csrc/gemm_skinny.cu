// Skinny GEMM for decode-sized problems (M <= 8 rows): C[M,N] = epi(A[M,K] . W[N,K]^T + bias) (+ residual).
//
// With one to eight activation rows every weight byte is used for <= 16 flops, so the problem is HBM-bound: the job is
// to stream W once at full bandwidth (Vicuna-7B decode reads 13.5 GB of weights per token; SURVEY.md §8d "secondary (HBM)
// bounds").  A 128-row tcgen05 tile wastes >= 94 % of the tensor core here and, worse, leaves most SMs without enough
// loads in flight (measured 0.15-0.46 of the HBM roofline on the decode shapes).  This kernel keeps the machine full of
// outstanding 16-byte loads instead:
//   * one warp owns two adjacent weight rows (= two output columns; for the interleaved gate/up matrix exactly one SwiGLU
//     output) and strides over K with 16 B per lane => every load instruction of a warp covers 512 contiguous bytes of a row;
//   * 8 such loads per row are issued back to back before the first FMA (8 KB in flight per warp), bypassing L1
//     allocation so the weights do not wash the activation rows out of L1;
//   * the M activation rows are re-read through L1 (they are at most 8 x 22 KB and shared by all warps of the SM);
//   * fp32 accumulation in the same k order for every column, butterfly reduction at the end, and the same epilogue
//     semantics as the tcgen05 kernel (bias -> activation / SwiGLU -> + residual -> bf16 / fp32, output-row remap).
// Algorithmic bytes per launch: 2*N*K (weights) + 2*M*K + 2*M*N.
#include "common.cuh"
#include "vstar_b200.h"

namespace {

struct SkinnyParams {
  const bf16* A;
  const bf16* W;
  void* C;
  const bf16* bias;
  const void* residual;
  long long lda, ldw, ldc, ldr;
  int M, N, K;
  int epilogue, out_fp32;
  int rows_per_group;
  long long group_stride, group_offset;
};

__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  // bf16 -> fp32 is a 16-bit shift: low half = element 0
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}

constexpr int SK_THREADS = 256;
constexpr int SK_UNROLL = 8;      // 16 B loads in flight per lane and weight row

template <int MT>
__global__ void __launch_bounds__(SK_THREADS) gemm_skinny_kernel(const SkinnyParams p) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * SK_THREADS + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * SK_THREADS) >> 5;
  const int npairs = (p.N + 1) >> 1;
  const int kchunks = p.K >> 3;                 // K % 8 == 0 (checked on the host)
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);

  for (int pair = warp; pair < npairs; pair += nwarps) {
    const int n0 = pair * 2;
    const bool has1 = (n0 + 1 < p.N);
    const uint4* w0 = reinterpret_cast<const uint4*>(p.W + (long long)n0 * p.ldw);
    const uint4* w1 = reinterpret_cast<const uint4*>(p.W + (long long)(has1 ? n0 + 1 : n0) * p.ldw);
    float acc[MT][2];
#pragma unroll
    for (int m = 0; m < MT; ++m) acc[m][0] = acc[m][1] = 0.f;

    for (int c0 = lane; c0 < kchunks; c0 += 32 * SK_UNROLL) {
      uint4 wa[SK_UNROLL], wb[SK_UNROLL];
#pragma unroll
      for (int u = 0; u < SK_UNROLL; ++u) {
        const int c = c0 + u * 32;
        const bool ok = c < kchunks;
        wa[u] = ok ? ld_stream16(w0 + c) : zero;
        wb[u] = ok ? ld_stream16(w1 + c) : zero;
      }
#pragma unroll
      for (int u = 0; u < SK_UNROLL; ++u) {
        const int c = c0 + u * 32;
        if (c < kchunks) {
          float fa[8], fb[8];
          unpack8(wa[u], fa);
          unpack8(wb[u], fb);
#pragma unroll
          for (int m = 0; m < MT; ++m) {
            if (m < p.M) {
              const uint4 av = __ldg(reinterpret_cast<const uint4*>(p.A + (long long)m * p.lda) + c);
              float x[8];
              unpack8(av, x);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                acc[m][0] = fmaf(x[i], fa[i], acc[m][0]);
                acc[m][1] = fmaf(x[i], fb[i], acc[m][1]);
              }
            }
          }
        }
      }
    }
    // butterfly: every lane ends with the full sums
#pragma unroll
    for (int m = 0; m < MT; ++m) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        acc[m][0] += __shfl_xor_sync(0xffffffffu, acc[m][0], o);
        acc[m][1] += __shfl_xor_sync(0xffffffffu, acc[m][1], o);
      }
    }
    // lane 2m + c finishes output (m, n0 + c)
    float mine = 0.f;
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      if (lane == 2 * m) mine = acc[m][0];
      if (lane == 2 * m + 1) mine = acc[m][1];
    }
    const int m = lane >> 1, c = lane & 1;
    const int n = n0 + c;
    const bool valid = (m < p.M) && (m < MT) && (n < p.N);
    if (valid && p.bias != nullptr) mine += bf2f(p.bias[n]);
    const float up = __shfl_down_sync(0xffffffffu, mine, 1);        // SwiGLU: odd column = up_j
    if (!valid) continue;
    const long long orow = (long long)(m / p.rows_per_group) * p.group_stride + p.group_offset + (m % p.rows_per_group);
    if (p.epilogue == VSB_EPI_SWIGLU) {
      if (c == 0) reinterpret_cast<bf16*>(p.C)[orow * p.ldc + (n0 >> 1)] = f2bf(silu_f(mine) * up);
      continue;
    }
    if (p.epilogue == VSB_EPI_QUICK_GELU) mine = quick_gelu_f(mine);
    else if (p.epilogue == VSB_EPI_RELU) mine = fmaxf(mine, 0.f);
    else if (p.epilogue == VSB_EPI_GELU) mine = gelu_erf_f(mine);
    if (p.out_fp32) {
      if (p.residual) mine += reinterpret_cast<const float*>(p.residual)[orow * p.ldr + n];
      reinterpret_cast<float*>(p.C)[orow * p.ldc + n] = mine;
    } else {
      if (p.residual) mine += bf2f(reinterpret_cast<const bf16*>(p.residual)[orow * p.ldr + n]);
      reinterpret_cast<bf16*>(p.C)[orow * p.ldc + n] = f2bf(mine);
    }
  }
}

}  // namespace

// called by vsb_gemm_bf16 (gemm_tcgen05.cu) for M <= 8, K % 8 == 0; arguments already validated there
int vsb_gemm_skinny_launch(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc, int M, int N, int K,
                           const void* bias, const void* residual, long long ldr, int epilogue, int out_fp32, int rows_per_group,
                           long long group_stride, long long group_offset, cudaStream_t stream) {
  SkinnyParams p;
  p.A = reinterpret_cast<const bf16*>(A);
  p.W = reinterpret_cast<const bf16*>(W);
  p.C = C;
  p.bias = reinterpret_cast<const bf16*>(bias);
  p.residual = residual;
  p.lda = lda; p.ldw = ldw; p.ldc = ldc; p.ldr = ldr;
  p.M = M; p.N = N; p.K = K;
  p.epilogue = epilogue;
  p.out_fp32 = out_fp32;
  p.rows_per_group = rows_per_group;
  p.group_stride = group_stride;
  p.group_offset = group_offset;
  const int npairs = (N + 1) / 2;
  const int warps_per_block = SK_THREADS / 32;
  int blocks = (npairs + warps_per_block - 1) / warps_per_block;
  const int cap = vsb_num_sms() * 8;                    // grid-stride beyond 8 resident blocks per SM
  if (blocks > cap) blocks = cap;
  if (M <= 1) gemm_skinny_kernel<1><<<blocks, SK_THREADS, 0, stream>>>(p);
  else if (M <= 2) gemm_skinny_kernel<2><<<blocks, SK_THREADS, 0, stream>>>(p);
  else if (M <= 4) gemm_skinny_kernel<4><<<blocks, SK_THREADS, 0, stream>>>(p);
  else gemm_skinny_kernel<8><<<blocks, SK_THREADS, 0, stream>>>(p);
  VSB_LAUNCH_CHECK();
  return VSB_OK;
}
