// Micro-benchmark: MUFU.EX2 issue rate and latency per SM sub-partition (B200).  nvcc -arch=sm_100a -O3 -o mufu mufu.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void mufu_rate(float* out, long long* clk, int iters) {
  float x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = -0.001f * (threadIdx.x + i);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x % 32 == 0) clk[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
}

// the softmax inner pattern: x = fma(v, s, -m); p = ex2(x); acc += p; pack pairs
template <int N>
__global__ void softmax_pattern(const float* in, float* out, long long* clk, int iters) {
  float v[N];
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = in[(threadIdx.x * N + i) % 1024];
  float m = in[threadIdx.x % 7], sl2 = 1.3f;
  float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  unsigned pkacc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < N; c += 4) {
      float p0, p1, p2, p3;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(v[c], sl2, -m)));
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(v[c + 1], sl2, -m)));
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p2) : "f"(fmaf(v[c + 2], sl2, -m)));
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p3) : "f"(fmaf(v[c + 3], sl2, -m)));
      a0 += p0; a1 += p1; a2 += p2; a3 += p3;
      unsigned w0, w1;
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w0) : "f"(p1), "f"(p0));
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w1) : "f"(p3), "f"(p2));
      pkacc ^= w0 ^ w1;
    }
    m += 1e-3f;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + __uint_as_float(pkacc);
  if (threadIdx.x % 32 == 0) clk[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
}

// same work, software-pipelined in chunks of 8 with the consumers of chunk k-1 tied behind chunk k's last exponential
template <int N>
__global__ void softmax_pattern_tied(const float* in, float* out, long long* clk, int iters) {
  float v[N];
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = in[(threadIdx.x * N + i) % 1024];
  float m = in[threadIdx.x % 7], sl2 = 1.3f;
  float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  unsigned pkacc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float pv[8];
#pragma unroll
    for (int c = 0; c < N + 8; c += 8) {
      float cur[8];
      if (c < N) {
#pragma unroll
        for (int t = 0; t < 8; ++t) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(cur[t]) : "f"(fmaf(v[c + t], sl2, -m)));
      }
      if (c > 0) {
        const float tie = (c < N) ? fmaf(cur[7], 0.0f, 1.0f) : 1.0f;
        float q[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) q[t] = pv[t] * tie;
        a0 += q[0] + q[4]; a1 += q[1] + q[5]; a2 += q[2] + q[6]; a3 += q[3] + q[7];
#pragma unroll
        for (int t = 0; t < 8; t += 2) {
          unsigned w;
          asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(q[t + 1]), "f"(q[t]));
          pkacc ^= w;
        }
      }
#pragma unroll
      for (int t = 0; t < 8; ++t) pv[t] = cur[t];
    }
    m += 1e-3f;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + __uint_as_float(pkacc);
  if (threadIdx.x % 32 == 0) clk[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
}

int main() {
  float *out, *in;
  long long* clk;
  cudaMalloc(&out, 1 << 20);
  cudaMalloc(&in, 4096);
  cudaMemset(in, 0, 4096);
  cudaMalloc(&clk, 4096);
  long long h[64];
  const int iters = 256;
  for (int warps_per_smsp = 1; warps_per_smsp <= 2; ++warps_per_smsp) {
    int threads = 128 * warps_per_smsp;
    mufu_rate<8><<<1, threads>>>(out, clk, iters);
    cudaMemcpy(h, clk, sizeof(long long) * threads / 32, cudaMemcpyDeviceToHost);
    printf("mufu_rate ILP=8  warps/SMSP=%d: %.2f clk per warp-wide ex2 (warp 0)\n", warps_per_smsp, (double)h[0] / (iters * 8));
    mufu_rate<1><<<1, threads>>>(out, clk, iters);
    cudaMemcpy(h, clk, sizeof(long long) * threads / 32, cudaMemcpyDeviceToHost);
    printf("mufu_rate ILP=1  warps/SMSP=%d: %.2f clk per dependent ex2 (latency)\n", warps_per_smsp, (double)h[0] / iters);
    softmax_pattern<64><<<1, threads>>>(in, out, clk, iters);
    cudaMemcpy(h, clk, sizeof(long long) * threads / 32, cudaMemcpyDeviceToHost);
    printf("softmax_pattern N=64  warps/SMSP=%d: %.2f clk per element\n", warps_per_smsp, (double)h[0] / (iters * 64));
    softmax_pattern_tied<64><<<1, threads>>>(in, out, clk, iters);
    cudaMemcpy(h, clk, sizeof(long long) * threads / 32, cudaMemcpyDeviceToHost);
    printf("softmax_pattern_tied N=64  warps/SMSP=%d: %.2f clk per element\n", warps_per_smsp, (double)h[0] / (iters * 64));
    softmax_pattern_tied<128><<<1, threads>>>(in, out, clk, iters);
    cudaMemcpy(h, clk, sizeof(long long) * threads / 32, cudaMemcpyDeviceToHost);
    printf("softmax_pattern_tied N=128 warps/SMSP=%d: %.2f clk per element\n", warps_per_smsp, (double)h[0] / (iters * 128));
    softmax_pattern<128><<<1, threads>>>(in, out, clk, iters);
    cudaMemcpy(h, clk, sizeof(long long) * threads / 32, cudaMemcpyDeviceToHost);
    printf("softmax_pattern N=128 warps/SMSP=%d: %.2f clk per element\n", warps_per_smsp, (double)h[0] / (iters * 128));
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
