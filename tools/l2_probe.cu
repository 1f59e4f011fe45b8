// L2 residency / per-SM L2 bandwidth probe (design input for the persistent window step).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/l2_probe tools/l2_probe.cu && tools/l2_probe
// Test 1: a working set of N MB is streamed cyclically (read + write in place, float4) by all SMs, R passes inside
//         one launch; GB/s per pass vs N shows where the set stops living in L2.
// Test 2: the same with n active CTAs (1 per SM) on a 32 MB set: L2 bandwidth one SM can pull.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

__global__ void __launch_bounds__(512) stream_rw(float4* __restrict__ buf, size_t n4, int passes, int write) {
  const size_t per = (n4 + gridDim.x - 1) / gridDim.x;
  const size_t lo = per * blockIdx.x, hi = (lo + per < n4) ? lo + per : n4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p = 0; p < passes; ++p) {
    for (size_t i = lo + threadIdx.x; i < hi; i += 4 * 512) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (i + u * 512 < hi) ? buf[i + u * 512] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (write) {
          v[u].x = v[u].x * 0.999f + 1e-3f; v[u].y += 1e-6f; v[u].z *= 1.0001f; v[u].w -= 1e-6f;
          if (i + u * 512 < hi) buf[i + u * 512] = v[u];
        } else {
          acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
        }
      }
    }
  }
  if (!write && acc.x + acc.y + acc.z + acc.w == 1234.5f) buf[lo] = acc;
}

static float run(float4* buf, size_t bytes, int ctas, int passes, int write) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  stream_rw<<<ctas, 512>>>(buf, bytes / 16, 2, write);  // warm
  cudaEventRecord(a);
  stream_rw<<<ctas, 512>>>(buf, bytes / 16, passes, write);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  return ms;
}

int main() {
  const size_t max_bytes = (size_t)512 << 20;
  float4* buf = nullptr;
  if (cudaMalloc(&buf, max_bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMemset(buf, 0, max_bytes);
  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
  printf("# test 1: cyclic working set, %d CTAs x 512 threads, 20 passes\n", n_sm);
  printf("%8s %6s %12s %12s\n", "MB", "mode", "us/pass", "GB/s(r+w)");
  const int sizes[] = {8, 16, 32, 48, 64, 80, 96, 112, 128, 160, 192, 256, 512};
  for (int s : sizes)
    for (int write = 0; write < 2; ++write) {
      const size_t bytes = (size_t)s << 20;
      const int passes = 20;
      const float ms = run(buf, bytes, n_sm, passes, write);
      const double per = ms * 1e3 / passes;
      printf("%8d %6s %12.2f %12.1f\n", s, write ? "rw" : "r", per, (write ? 2.0 : 1.0) * bytes / (per * 1e-6) / 1e9);
    }
  printf("# test 2: n CTAs on a 32 MB set (L2 resident), 20 passes\n");
  printf("%8s %6s %12s %14s\n", "CTAs", "mode", "GB/s total", "GB/s per CTA");
  const int ns[] = {1, 2, 8, 16, 37, 74, 148, 296};
  for (int n : ns)
    for (int write = 0; write < 2; ++write) {
      const size_t bytes = (size_t)32 << 20;
      const int passes = n < 8 ? 4 : 20;
      const float ms = run(buf, bytes, n, passes, write);
      const double gbs = (write ? 2.0 : 1.0) * bytes * passes / (ms * 1e-3) / 1e9;
      printf("%8d %6s %12.1f %14.2f\n", n, write ? "rw" : "r", gbs, gbs / n);
    }
  cudaFree(buf);
  return 0;
}
