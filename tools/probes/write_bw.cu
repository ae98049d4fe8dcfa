// tools/probes/write_bw.cu -- what is the HBM roofline of a WRITE stream on this GPU?
// The fused traversal kernels write every CLV once (0.96 GB per evaluation at 100 taxa x 100k sites) and read
// almost nothing from DRAM (children are forwarded on chip or still in L2), so their memory roofline is the
// bandwidth of a pure write stream, not the read+write figure of a copy.  This probe measures both, with
// the same 256-bit stores the kernels use.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o write_bw write_bw.cu && ./write_bw
#include <cstdio>
#include <cuda_runtime.h>

struct __align__(32) d4 { double x, y, z, w; };

__global__ void k_write256(d4 *p, size_t n, double v)
{
  d4 val{v, v, v, v};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p + i), "d"(val.x), "d"(val.y), "d"(val.z), "d"(val.w) : "memory");
}
__global__ void k_write256_cs(d4 *p, size_t n, double v)
{
  d4 val{v, v, v, v};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p + i), "d"(val.x), "d"(val.y), "d"(val.z), "d"(val.w) : "memory");
}
__global__ void k_write128(double2 *p, size_t n, double v)
{
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = make_double2(v, v);
}
__global__ void k_copy256(const d4 *__restrict__ a, d4 *__restrict__ b, size_t n)
{
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
  {
    d4 v;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(a + i) : "memory");
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(b + i), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
  }
}
__global__ void k_read256(const d4 *__restrict__ a, size_t n, double *out)
{
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
  {
    d4 v;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(a + i) : "memory");
    s += v.x + v.y + v.z + v.w;
  }
  if (s == 12345.678) *out = s;
}
// the traversal's pattern: per "update" write a fresh 10 MB buffer and read the one written two updates ago (L2 hit)
__global__ void k_write_read_recent(d4 *base, size_t per_buf, int n_buf, double *out)
{
  double s = 0.0;
  for (int b = 0; b < n_buf; ++b)
  {
    d4       *dst = base + (size_t)b * per_buf;
    const d4 *src = base + (size_t)(b >= 2 ? b - 2 : 0) * per_buf;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_buf; i += (size_t)gridDim.x * blockDim.x)
    {
      d4 v{1.0, 1.0, 1.0, 1.0};
      if (b >= 2) asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(src + i) : "memory");
      v.x += 1.0;
      asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(dst + i), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
      s += v.y;
    }
  }
  if (s == 12345.678) *out = s;
}

template <typename F>
static float time_it(F f, int reps = 10)
{
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  f();
  f();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r)
  {
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}

int main()
{
  const size_t bytes = (size_t)4 << 30;
  char        *a, *b;
  double      *out;
  cudaMalloc(&a, bytes);
  cudaMalloc(&b, bytes);
  cudaMalloc(&out, 8);
  cudaMemset(a, 1, bytes);
  cudaMemset(b, 1, bytes);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int grid = prop.multiProcessorCount * 8;
  const size_t n256 = bytes / 32, n128 = bytes / 16;
  float ms;
  ms = time_it([&] { cudaMemsetAsync(a, 0, bytes); });
  printf("cudaMemset            4 GiB: %7.3f ms  %7.1f GB/s written\n", ms, bytes / ms / 1e6);
  ms = time_it([&] { k_write256<<<grid, 512>>>((d4 *)a, n256, 1.0); });
  printf("st.global.v4.f64      4 GiB: %7.3f ms  %7.1f GB/s written\n", ms, bytes / ms / 1e6);
  ms = time_it([&] { k_write256_cs<<<grid, 512>>>((d4 *)a, n256, 1.0); });
  printf("st.global.cs.v4.f64   4 GiB: %7.3f ms  %7.1f GB/s written\n", ms, bytes / ms / 1e6);
  ms = time_it([&] { k_write128<<<grid, 512>>>((double2 *)a, n128, 1.0); });
  printf("st.global.v2.f64      4 GiB: %7.3f ms  %7.1f GB/s written\n", ms, bytes / ms / 1e6);
  ms = time_it([&] { k_read256<<<grid, 512>>>((const d4 *)a, n256, out); });
  printf("ld.global.v4.f64      4 GiB: %7.3f ms  %7.1f GB/s read\n", ms, bytes / ms / 1e6);
  ms = time_it([&] { k_copy256<<<grid, 512>>>((const d4 *)a, (d4 *)b, n256); });
  printf("copy 256-bit      4+4 GiB: %7.3f ms  %7.1f GB/s read+written (%.1f each way)\n", ms, 2.0 * bytes / ms / 1e6, bytes / ms / 1e6);
  ms = time_it([&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); });
  printf("cudaMemcpy D2D    4+4 GiB: %7.3f ms  %7.1f GB/s read+written\n", ms, 2.0 * bytes / ms / 1e6);
  {
    const size_t per_buf = (size_t)10 * 1000 * 1000 / 32;  // 10 MB like one CLV of 78 483 patterns
    const int    n_buf = 98;
    ms = time_it([&] { k_write_read_recent<<<prop.multiProcessorCount * 2, 512>>>((d4 *)a, per_buf, n_buf, out); });
    printf("98 x (write 10 MB, read the buffer written 2 steps earlier from L2): %7.3f ms  %7.1f GB/s written\n", ms,
           (double)per_buf * 32 * n_buf / ms / 1e6);
  }
  return 0;
}
